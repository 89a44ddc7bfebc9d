"""Load the UNMODIFIED reference modules (Python 2 sources) under Python 3.

TEST INFRASTRUCTURE ONLY.  This file is used by ``make_golden.py`` (and by the
optional live cross-checks in ``tests/test_oracle_vs_reference.py``) to execute
the reference's own functions from ``/root/reference`` in THIS container so
that golden input/output vectors can be generated and the oracle restatement
(``oracle/``) can be pinned against them.  Nothing here is importable from the
product package and nothing here runs on the GPU box (``/root/reference`` does
not exist there).

The reference (pennucci/PulsePortraiture) is Python-2 syntax and imports
``psrchive``/``matplotlib``/``lmfit``/``pywt`` unconditionally, so a plain
``import pplib`` fails.  The loader below reads the source text where it lies
(never copied into this repo), applies purely syntactic py2->py3 rewrites
(print statements, integer division where an index is built, removed numpy
aliases, lazy ``map``), stubs the absent third-party modules, and ``exec``s the
result top-level-block by top-level-block so that blocks which still do not
compile (PSRCHIVE plumbing, plotting) are skipped without affecting the
numerical functions (SURVEY.md section 8c documents the recipe).
"""
from __future__ import annotations

import os
import re
import sys
import types

REFERENCE_DIR = os.environ.get("PP_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "pplib.py"))


# --------------------------------------------------------------------------
# stubs for absent third-party modules
# --------------------------------------------------------------------------
class _Anything(types.ModuleType):
    """Module stub: any attribute is another stub; calling returns a stub."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")


def _install_stubs():
    for name in ("psrchive", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.gridspec", "matplotlib.patches",
                 "matplotlib.widgets", "lmfit", "pywt"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    # `import matplotlib.gridspec as gs` needs attribute access on the parent
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "gridspec", "patches", "widgets"):
        setattr(mpl, sub, sys.modules["matplotlib." + sub])


# --------------------------------------------------------------------------
# py2 -> py3 text rewrites
# --------------------------------------------------------------------------
def _bracket_balance(s: str) -> int:
    """Net open brackets in s, ignoring string literals (good enough here)."""
    s = re.sub(r'"(\\.|[^"\\])*"', '""', s)
    s = re.sub(r"'(\\.|[^'\\])*'", "''", s)
    s = s.split("#", 1)[0]
    return sum(s.count(c) for c in "([{") - sum(s.count(c) for c in ")]}")


_PRINT_RE = re.compile(r"^(?P<head>\s*(?:.*?:\s+)?)print(?:\s+(?P<body>.*))?$")


def _convert_prints(lines):
    out = []
    i = 0
    n = len(lines)
    while i < n:
        line = lines[i]
        stripped = line.lstrip()
        m = None
        if re.match(r"^\s*print(\s|$)", line) or re.search(r":\s+print\s", line):
            if not stripped.startswith("#"):
                m = _PRINT_RE.match(line.rstrip("\n"))
        if m is None or (m.group("head").strip() and
                         not m.group("head").rstrip().endswith(":")):
            out.append(line)
            i += 1
            continue
        head = m.group("head")
        body = m.group("body") or ""
        # gather continuation lines
        buf = body
        while (buf.rstrip().endswith("\\") or _bracket_balance(buf) > 0) \
                and i + 1 < n:
            i += 1
            nxt = lines[i].rstrip("\n")
            if buf.rstrip().endswith("\\"):
                buf = buf.rstrip()[:-1] + " " + nxt.strip()
            else:
                buf = buf + " " + nxt.strip()
        buf = buf.strip()
        end = ""
        if buf.endswith(","):
            buf = buf[:-1]
            end = ", end=' '"
        if buf.startswith(">>"):
            # print >>f, x  ->  print(x, file=f)
            tgt, _, rest = buf[2:].partition(",")
            out.append("%sprint(%s, file=%s)\n" % (head, rest.strip(), tgt.strip()))
        else:
            out.append("%sprint(%s%s)\n" % (head, buf, end))
        i += 1
    return out


_SUBS = [
    (r"\.itervalues\(\)", ".values()"),
    (r"\.iteritems\(\)", ".items()"),
    (r"(\w+)\.has_key\(([^)]*)\)", r"(\2 in \1)"),
    (r"\bxrange\(", "range("),
    # integer division where an index / count is built
    (r"nharm = nbin/2 \+ 1", "nharm = nbin//2 + 1"),
    (r"ngauss = \(len\(params\) - 2\) / 3", "ngauss = (len(params) - 2) // 3"),
    (r"ngauss = \(len\(([^)]*)\) - 2\) / 6", r"ngauss = (len(\1) - 2) // 6"),
    (r"ngauss = \(len\(([^)]*)\) - 2\) / 3", r"ngauss = (len(\1) - 2) // 3"),
    (r"mid = repeat/2", "mid = repeat//2"),
    (r"arr\.size/2 \+ 1", "arr.size//2 + 1"),
    # removed numpy aliases
    (r"dtype='complex_'", "dtype=complex"),
    (r"\bnp\.float\(", "float("),
    (r"\bnp\.bool\(", "bool("),
    (r"\bnp\.int\(", "int("),
    (r"\bnp\.float\b(?!\d|_)", "float"),
    (r"\bnp\.bool\b(?!_)", "bool"),
    (r"\bnp\.int\b(?!\d|_|e)", "int"),
    # lazy map / range
    (r"map\(bool, fit_flags\)", "list(map(bool, fit_flags))"),
    (r"comp = map\(np\.float64,", "comp = list(map(np.float64,"),
    (r"fit_comp = map\(int,", "fit_comp = list(map(int,"),
    (r"iaxis = range\(ndim\)", "iaxis = list(range(ndim))"),
]


def _fix_map_closers(text: str) -> str:
    # the two read_model map(...) rewrites above opened one extra paren each
    text = re.sub(r"(comp = list\(map\(np\.float64, [^\n]*\))", r"\1)", text)
    text = re.sub(r"(fit_comp = list\(map\(int, [^\n]*\))", r"\1)", text)
    return text


def _translate(src: str) -> str:
    lines = src.splitlines(keepends=True)
    lines = _convert_prints(lines)
    text = "".join(lines)
    for pat, rep in _SUBS:
        text = re.sub(pat, rep, text)
    text = _fix_map_closers(text)
    return text


def _top_level_blocks(text: str):
    """Split translated source into top-level statements (by indentation)."""
    lines = text.splitlines(keepends=True)
    blocks, cur = [], []
    in_triple = None
    depth = 0
    for line in lines:
        starts_block = False
        if in_triple is None and depth <= 0 and cur:
            s = line
            if s[:1] not in (" ", "\t", "\n", "\r", "#", ")", "]", "}") and s.strip():
                # decorators / else / elif / except / finally belong to previous
                if not re.match(r"^(else|elif|except|finally)\b", s):
                    starts_block = True
        if starts_block:
            blocks.append("".join(cur))
            cur = []
            depth = 0
        cur.append(line)
        # track triple-quoted strings crudely
        tmp = line
        while True:
            if in_triple is None:
                m = re.search(r'("""|\'\'\')', tmp)
                if not m:
                    break
                in_triple = m.group(1)
                tmp = tmp[m.end():]
            else:
                idx = tmp.find(in_triple)
                if idx < 0:
                    break
                tmp = tmp[idx + 3:]
                in_triple = None
        if in_triple is None:
            depth += _bracket_balance(line)
    if cur:
        blocks.append("".join(cur))
    return blocks


_CACHE = {}


def load_reference(verbose: bool = False):
    """Return (pplib, pptoaslib) module objects exec'd from the reference."""
    if "mods" in _CACHE:
        return _CACHE["mods"]
    if not reference_available():
        raise RuntimeError("reference not found under %s" % REFERENCE_DIR)
    _install_stubs()
    tc = types.ModuleType("telescope_codes")
    tc.telescope_code_dict = {}
    sys.modules.setdefault("telescope_codes", tc)

    mods = {}
    for name in ("pplib", "pptoaslib"):
        path = os.path.join(REFERENCE_DIR, name + ".py")
        with open(path, "r") as fh:
            text = _translate(fh.read())
        mod = types.ModuleType("ref_" + name)
        mod.__file__ = path
        if name == "pptoaslib":
            # `from pplib import *`
            mod.__dict__.update({k: v for k, v in mods["pplib"].__dict__.items()
                                 if not k.startswith("__")})
            text = text.replace("from pplib import *", "pass")
        ok = bad = 0
        for blk in _top_level_blocks(text):
            if not blk.strip():
                continue
            try:
                code = compile(blk, path, "exec")
                exec(code, mod.__dict__)
                ok += 1
            except Exception as exc:  # noqa: BLE001 - skip PSRCHIVE/plot blocks
                bad += 1
                if verbose:
                    first = blk.strip().splitlines()[0][:70]
                    print("[ref_shim] skipped block in %s: %s (%s: %s)" % (
                        name, first, type(exc).__name__, str(exc)[:80]))
        if verbose:
            print("[ref_shim] %s: %d blocks loaded, %d skipped" % (name, ok, bad))
        mods[name] = mod
    _CACHE["mods"] = (mods["pplib"], mods["pptoaslib"])
    return _CACHE["mods"]


if __name__ == "__main__":
    pl, ptl = load_reference(verbose=True)
    need = ["fit_portrait", "fit_phase_shift", "get_noise", "rotate_data",
            "rotate_portrait", "phase_transform", "guess_fit_freq",
            "read_model", "gen_gaussian_portrait", "get_bin_centers",
            "scattering_portrait_FT", "scattering_times", "get_scales"]
    for fn in need:
        assert hasattr(pl, fn), fn
    for fn in ["fit_portrait_full", "get_nu_zeros", "phase_shifts",
               "fit_portrait_full_function_2deriv_with_scales"]:
        assert hasattr(ptl, fn), fn
    print("reference functions loaded OK")

"""Golden .tim lines for write_TOAs (pplib.py:3445-3503): run the reference's own
function (through ref_shim) on a fixed list of TOAs and store its output next to a JSON
description of the inputs.  TEST INFRASTRUCTURE; needs /root/reference.

    python tests/golden/make_golden_toas.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import ref_shim  # noqa: E402

mods = ref_shim.load_reference()
pl = mods["pplib"] if isinstance(mods, dict) else mods[0]


class FakeMJD(object):
    def __init__(self, day, frac):
        self.day, self.frac = day, frac

    def intday(self):
        return self.day

    def fracday(self):
        return self.frac


class FakeTOA(object):
    def __init__(self, spec):
        self.archive = spec["archive"]
        self.frequency = np.inf if spec["frequency"] == "inf" else spec["frequency"]
        self.MJD = FakeMJD(spec["mjd"][0], spec["mjd"][1])
        self.TOA_error = spec["err"]
        self.telescope, self.telescope_code = spec["telescope"], spec["code"]
        self.DM, self.DM_error = spec["DM"], spec["DM_error"]
        self.flags = dict(spec["flags"])
        for k, v in self.flags.items():
            setattr(self, k, v)


SPECS = [
    dict(archive="guppi_56000_J1234.fits", frequency=1441.12345678901, mjd=[56000, 0.123456789012345], err=0.4321,
         telescope="GBT", code="1", DM=15.9876543, DM_error=0.0001234,
         flags=[["be", "GUPPI"], ["fe", "Rcvr1_2"], ["f", "Rcvr1_2_GUPPI"], ["nbin", 2048], ["nch", 512], ["nchx", 500],
                ["bw", 783.125], ["chbw", 1.5625], ["subint", 3], ["tobs", 119.8765], ["fratio", 1.62345],
                ["tmplt", "J1234.spl"], ["snr", 234.56789], ["gof", 1.02345], ["phi_DM_cov", -1.234e-9],
                ["phs", 0.1234567891], ["phs_err", 1.23e-5], ["flux", 1.234567], ["flux_err", 0.0123456]]),
    dict(archive="puppi_57000_J0000.fits", frequency="inf", mjd=[57000, 0.999999999999999], err=12.0,
         telescope="Arecibo", code="ao", DM=None, DM_error=None,
         flags=[["be", "PUPPI"], ["subint", 0], ["snr", 9.5], ["gof", 0.98]]),
    dict(archive="low_snr.fits", frequency=820.5, mjd=[55555, 0.5], err=100.0, telescope="GBT", code="1",
         DM=10.0, DM_error=0.1, flags=[["snr", 3.0]]),                    # cut by SNR_cutoff = 8
    dict(archive="no_snr_flag.fits", frequency=820.5, mjd=[55555, 0.25], err=1.0, telescope="GBT", code="1",
         DM=None, DM_error=None, flags=[["subint", 1]]),                  # no snr attribute: culled
]

# Python 2's exec statement rebinds function locals, Python 3's exec() cannot: re-create
# write_TOAs from the reference's text with  exec("toa_string += '...'"%(...))  turned into the
# statement it executes (purely syntactic; same shim rules as ref_shim otherwise).
import re  # noqa: E402
src = open(os.path.join(ref_shim.REFERENCE_DIR, "pplib.py")).read()
beg = src.index("def write_TOAs(")
end = src.index("\ndef ", beg + 10)
fn = src[beg:end]
fn = re.sub(r'exec\("toa_string \+= (\'[^\']*\')"%\((.*?)\)\)', r"toa_string += \1%(\2)", fn, flags=re.S)
fn = fn.replace(".iteritems()", ".items()").replace("print toa_string", "print(toa_string)")
assert "exec(" not in fn
ns = dict(vars(pl)) if not isinstance(pl, dict) else dict(pl)
exec(compile(fn, "write_TOAs(reference text)", "exec"), ns)
pl_write_TOAs = ns["write_TOAs"]

out = os.path.join(HERE, "toas_v1.tim")
if os.path.exists(out):
    os.remove(out)
pl_write_TOAs([FakeTOA(s) for s in SPECS], inf_is_zero=True, SNR_cutoff=8.0, outfile=out, append=False)
pl_write_TOAs([FakeTOA(SPECS[1])], inf_is_zero=False, SNR_cutoff=0.0, outfile=out, append=True)
json.dump(SPECS, open(os.path.join(HERE, "toas_v1.json"), "w"), indent=1)
print(open(out).read())

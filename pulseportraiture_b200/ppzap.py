"""``ppzap.get_zap_channels`` (ppzap.py:18-48): median/sigma iteration on the
per-channel noise levels (thin host logic over quantities the fit already has;
the noise levels themselves come from the device, pp_get_noise_batch)."""
from __future__ import annotations

import numpy as np


def get_zap_channels(data, nstd=3):
    """Return, per good subint, the channels whose noise level exceeds the
    subint median by more than nstd standard deviations, iterated until no
    channel is flagged (ppzap.py:18-48)."""
    zap_channels = []
    for isub in data.ok_isubs:
        ichans = np.array(data.ok_ichans[isub], dtype=int)
        zapped = []
        while len(ichans):
            noise = np.asarray(data.noise_stds)[isub, 0, ichans]
            bad = noise > np.median(noise) + nstd * np.std(noise)
            if not bad.any():
                break
            zapped.extend(ichans[bad].tolist())
            ichans = ichans[~bad]
        zap_channels.append(sorted(zapped))
    return zap_channels


def print_paz_cmds(datafiles, zap_channels, all_subs=False, modify=True, outfile=None, quiet=False):
    """paz command lines for the proposed channels (ppzap.py:50-96)."""
    lines = []
    for datafile, per_sub in zip(datafiles, zap_channels):
        if all_subs:
            chans = sorted(set(c for sub in per_sub for c in sub))
            if chans:
                lines.append("paz %s -z '%s' %s" % ("-m" if modify else "-e zap",
                                                     " ".join(map(str, chans)), datafile))
        else:
            for isub, chans in enumerate(per_sub):
                for c in chans:
                    lines.append("paz %s -I -z %d -w %d %s" % ("-m" if modify else "-e zap", c,
                                                               isub, datafile))
    text = "\n".join(lines)
    if outfile is not None:
        open(outfile, "a").write(text + ("\n" if text else ""))
    elif not quiet:
        print(text)
    return lines

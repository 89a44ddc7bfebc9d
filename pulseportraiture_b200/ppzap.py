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
    """paz command lines for the proposed channels (ppzap.py:50-96).

    modify=True edits the archive in place (``paz -m ... <datafile>``); otherwise one
    ``paz -e zap <datafile>`` writes <datafile stem>.zap and every following command modifies that
    copy.  all_subs=True zaps a channel flagged in any subint in all of them.  The lines go to
    standard output, or are appended to ``outfile``; they are also returned."""
    if not len(datafiles) or not len(zap_channels):
        if not quiet:
            print("Nothing to zap.")
        return None
    lines = []
    for iarch, datafile in enumerate(datafiles):
        count = sum(len(sub) for sub in zap_channels[iarch])
        paz_outfile = datafile
        if count and not modify:
            ii = datafile[::-1].find(".")
            paz_outfile = datafile + ".zap" if ii < 0 else datafile[:-ii] + "zap"
            lines.append("paz -e zap %s" % datafile)
        last_line = ""
        for isub, bad_ichans in enumerate(zap_channels[iarch]):
            for bad_ichan in bad_ichans:
                if not all_subs:
                    lines.append("paz -m -I -z %d -w %d %s" % (bad_ichan, isub, paz_outfile))
                else:
                    line = "paz -m -z %d %s" % (bad_ichan, paz_outfile)
                    if line != last_line:
                        lines.append(line)
                    last_line = line
    text = "\n".join(lines)
    if outfile is not None:
        with open(outfile, "a") as fh:
            fh.write(text + ("\n" if text else ""))
        if not quiet:
            print("Wrote %s." % outfile)
    else:
        print(text)
    return lines

"""Extracts the reference's own saved tracking runs into tests/golden/scilab_track_golden.npz (and ..._l2.npz):
SCI/GLONASS/L1/trackingResults.dat and SCI/GLONASS/L2/trackingResults.dat (Scilab 5 save() of trackResults, settings,
acqResults, channel written by postProcessing.sce:143 after tracking 1500 ms of a real GLONASS recording on the L1 and
on the L2 front-end settings), read with oracle/scilab_save.py.
Run in the container that has /root/reference:   python tests/golden/make_scilab_track_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scilab_save  # noqa: E402

SRC = "/root/reference/trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_SCILAB_RECEIVERS/GLONASS/L1/trackingResults.dat"
SRC_L2 = SRC.replace("/L1/", "/L2/")


def extract(path=SRC):
    v = scilab_save.load(path)
    track = v[[k for k in v if k.lower().startswith("track")][0]]  # the name field of the file reads "trackRdsults"
    out = {}
    for f in ("I_E", "I_P", "I_L", "Q_E", "Q_P", "Q_L", "carrFreq", "codeFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt",
              "absoluteSample"):
        out["track_" + f] = np.asarray(track[f][0], dtype=np.float64).ravel()  # channel 1 of 2 (the second one is empty)
    out["track_status"] = np.array([str(x.ravel()[0]) for x in track["status"]])
    for f in ("msToProcess", "numberOfChannels", "skipNumberOfBytes", "fileType", "samplingFreq", "codeFreqBasis", "IF", "L1_IF_step",
              "codeLength", "dllDampingRatio", "dllNoiseBandwidth", "dllCorrelatorSpacing", "pllDampingRatio", "pllNoiseBandwidth",
              "fllNoiseBandwidth"):
        out["settings_" + f] = np.float64(np.ravel(v["settings"][f])[0])
    for f in ("carrFreq", "codePhase", "peakMetric", "freqChannel"):
        out["acq_" + f] = np.asarray(v["acqResults"][f], dtype=np.float64).ravel()
    for f in ("SVN", "FCH", "acquiredFreq", "codePhase"):
        out["channel_" + f] = np.array([np.ravel(x)[0] if np.size(x) else 0.0 for x in v["channel"][f]], dtype=np.float64)
    return out


if __name__ == "__main__":
    for src, name in ((SRC, "scilab_track_golden.npz"), (SRC_L2, "scilab_track_golden_l2.npz")):
        o = extract(src)
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name), **o)
        print(name, "written", len(o), "arrays;", int(o["settings_msToProcess"]), "ms, FCH", o["channel_FCH"])

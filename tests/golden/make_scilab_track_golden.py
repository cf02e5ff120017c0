"""Extracts the reference's own saved tracking run into tests/golden/scilab_track_golden.npz:
SCI/GLONASS/L1/trackingResults.dat (Scilab 5 save() of trackResults, settings, acqResults, channel written by
postProcessing.sce:143 after tracking 1500 ms of a real GLONASS L1 recording), read with oracle/scilab_save.py.
Run in the container that has /root/reference:   python tests/golden/make_scilab_track_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scilab_save  # noqa: E402

SRC = "/root/reference/trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_SCILAB_RECEIVERS/GLONASS/L1/trackingResults.dat"


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
    o = extract()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "scilab_track_golden.npz"), **o)
    print("written", len(o), "arrays;", int(o["settings_msToProcess"]), "ms, FCH", o["channel_FCH"])

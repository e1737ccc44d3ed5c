#!/usr/bin/env python
"""Per-frame divergence of the adaptive state between the CUDA path and the NumPy oracle for chain B
(||dW||, ||dPhi_vv^-1||, dp reported per frame), from the waveform (device fp32 STFT) and
from the oracle's own spectrum.  Needs a GPU.  Prints a table; the committed copy is profiles/divergence_r01.txt."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import np_oracle as O                     # checker
    from distantspeech_b200.noise_estimation.mcspp_base import McSppBase
    from distantspeech_b200.transform.transform import Transform
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 256 * 250, seed0=0x5EED)[0].T)        # 4 s, [N, 8]
    taps = {}
    O.mvdr_mcspp_chain(x.astype(np.float64), geo, (30, 0), 512, 256, taps=taps)
    a0 = taps["a0"]
    for label, D in (("oracle spectrum in", taps["X"]),
                     ("waveform in (fp32 device STFT)", Transform(n_fft=512, hop_length=256, channel=8).stft(x))):
        est = McSppBase(nfft=512, channels=8)
        res = est.estimation_frames(D, a0=a0)
        T = D.shape[1]
        print("# %s: %d frames, 8 mics, 257 bins" % (label, T))
        print("# frame   max|dp|     max|dq|   rel||dxi||  rel||dgamma||  rel||dw_mvdr||   max|dG|")
        rows = []
        for n in range(T):
            rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
            rows.append((n, float(np.max(np.abs(res["p"][:, n] - taps["p"][:, n]))),
                         float(np.max(np.abs(res["q"][:, n] - taps["q"][:, n]))),
                         rel(res["xi"][:, n], taps["xi"][:, n]), rel(res["gamma"][:, n], taps["gamma"][:, n]),
                         rel(res["w_mvdr"][2:, n], taps["w"][2:, n]), float(np.max(np.abs(res["G"][:, n] - taps["G"][:, n])))))
        for r in rows:
            if r[0] < 4 or r[0] % 25 == 0 or r[0] == T - 1:
                print("%6d  %9.2e  %9.2e  %9.2e  %9.2e  %9.2e  %9.2e" % r)
        worst = np.max(np.array(rows)[:, 1:], axis=0)
        print("# worst over all frames: dp %.2e dq %.2e xi %.2e gamma %.2e w %.2e dG %.2e" % tuple(worst))
        Ai = est.Phi_vv_inv.real
        Ar = taps["est"].Phi_vv_inv.real
        print("# last frame: rel||dPhi_vv^-1|| = %.2e, rel||dPhi_vv|| = %.2e, rel||dPhi_yy|| = %.2e\n" % (
            np.linalg.norm(Ai - Ar) / np.linalg.norm(Ar),
            np.linalg.norm(est.Phi_vv - taps["est"].Phi_vv) / np.linalg.norm(taps["est"].Phi_vv),
            np.linalg.norm(est.Phi_yy - taps["est"].Phi_yy) / np.linalg.norm(taps["est"].Phi_yy)))


if __name__ == "__main__":
    main()

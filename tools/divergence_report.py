#!/usr/bin/env python
"""Per-frame divergence of the adaptive state between the CUDA path and the NumPy oracle (BASELINE.json north_star:
"adaptive NLMS/RLS state divergence reported per frame"; SURVEY.md 8d: ||dW||, ||dPhi_vv^-1||, dp per frame):

  * chain B (McSppBase + MVDR + OMLSA): p, q, xi, gamma, w, G per frame, from the waveform (device fp32 STFT) and from
    the oracle's own spectrum;
  * FDGSC (beamformer/FDGSC.py:201-317): rel ||dW_bm|| (all M blocking filters, gsc_bm.py:89-111) and rel ||dW_aic||
    (gsc_aic.py:81-97) per 256-sample block, fp32 and fp64 kernels;
  * SubbandLMS (adaptivefilter/SubbandLMS.py:28-84): rel ||dW||, rel ||dP|| per frame;
  * SubbandRLS (adaptivefilter/SubbandRLS.py:44-71): rel ||dW||, rel ||dP|| (inverse correlation matrices) per frame.

The device pipelines are stepped one block at a time (their state carries over between calls, bit-identical to one long
call) and the state is read back after every block.  Needs a GPU.  Prints tables; the committed copies are
profiles/divergence_r01.txt (chain B) and profiles/divergence_r02.txt (all four)."""
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


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def _table(title, header, rows, every):
    print("# " + title)
    print("# " + header)
    T = len(rows)
    for r in rows:
        if r[0] < 4 or r[0] % every == 0 or r[0] == T - 1:
            print(("%6d" + "  %9.2e" * (len(r) - 1)) % tuple(r))
    worst = np.max(np.array(rows)[:, 1:], axis=0)
    print("# worst over all %d steps: %s\n" % (T, "  ".join("%.2e" % w for w in worst)))


def fdgsc_report(nblk=250):
    from oracle import np_oracle as O
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.FDGSC import FDGSC
    geo = O.MicGeometry("linear", r=0.05, M=6, n_fft=256)
    ang = [60, 0]
    x = np.ascontiguousarray(O.synth_streams(1, geo, 256 * nblk, look_deg=(60.0, 0.0), interf_deg=(140.0, 0.0), seed0=33)[0].T)
    mic = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
    for prec in ("fp32", "fp64"):
        ref = O.FdgscOracle(geo, 256, np.array(ang) / 180 * np.pi)
        dev = FDGSC(mic, frameLen=256, angle=ang, precision=prec)
        rows = []
        for n in range(nblk):
            blk = x[n * 256:(n + 1) * 256]
            yo = ref.process(blk.astype(np.float64))[0]
            r = dev.process(blk.copy())
            Wbm_o = np.stack([f.W[:, 0] for f in ref.bm])
            Wbm_d = np.stack([f.W[:, 0] for f in dev.bm])
            rows.append((n, _rel(Wbm_d, Wbm_o), _rel(dev.aic_filter.W, ref.aic.W), float(np.max(np.abs(r[0] - yo))),
                         float(np.max(np.abs(r[1][:, 0] - ref.spp.p)))))
        _table("FDGSC (%s kernel), 6-mic linear, %d blocks of 256 samples" % (prec, nblk),
               " block  rel||dW_bm||  rel||dW_aic||  max|dy|   max|dp_mcra|", rows, 25)


def subband_report(nfrm=250):
    from oracle import np_oracle as O
    from distantspeech_b200.adaptivefilter.SubbandLMS import SubbandLMS
    from distantspeech_b200.adaptivefilter.SubbandRLS import SubbandRLS
    rng = np.random.default_rng(0x715)
    x = (rng.standard_normal(256 * nfrm) * 0.2).astype(np.float32)
    d = (0.5 * np.roll(x, 5) + 0.05 * rng.standard_normal(256 * nfrm)).astype(np.float32)
    K = 257
    ref, dev, rows = O.SubbandNlms(filter_len=2, num_bands=512, channel=1, mu=0.1, alpha=0.9), SubbandLMS(filter_len=2, num_bands=512), []
    for n in range(nfrm):
        xb, db = x[256 * n:256 * (n + 1)], d[256 * n:256 * (n + 1)]
        eo = ref.update(xb.astype(np.float64), db.astype(np.float64), np.ones(K))
        ed, _ = dev.update(xb.astype(np.float64), db.astype(np.float64), p=1.0)
        rows.append((n, _rel(dev.W, ref.W[:, :, 0]), _rel(dev.P, ref.P), float(np.max(np.abs(ed - eo)))))
    _table("SubbandLMS (NLMS, 2 frame taps, 512 bands, mu 0.1), %d frames" % nfrm, " frame  rel||dW||   rel||dP||   max|derr|", rows, 25)
    ref, dev, rows = O.SubbandRls(filter_len=2, num_bands=512), SubbandRLS(filter_len=2, num_bands=512), []
    for n in range(nfrm):
        xb, db = x[256 * n:256 * (n + 1)], d[256 * n:256 * (n + 1)]
        eo = ref.update(xb.astype(np.float64), db.astype(np.float64))
        ed, _ = dev.update(xb.astype(np.float64), db.astype(np.float64))
        rows.append((n, _rel(dev.W, ref.W), _rel(dev.P, ref.P), float(np.max(np.abs(ed - eo)))))
    _table("SubbandRLS (2 frame taps, 512 bands, lambda 0.998, mu 0.5), %d frames" % nfrm,
           " frame  rel||dW||   rel||dP||   max|derr|", rows, 25)


if __name__ == "__main__":
    what = sys.argv[1:] or ["chain", "fdgsc", "subband"]
    if "chain" in what:
        main()
    if "fdgsc" in what:
        fdgsc_report()
    if "subband" in what:
        subband_report()

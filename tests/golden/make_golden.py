"""Generate the golden fixtures in tests/golden by running the UNMODIFIED
reference (imported from /root/reference through oracle/ref_harness.py).

    python tests/golden/make_golden.py

Only runs in the build container (the reference tree is not on the GPU box);
the resulting .npz files are committed.  Inputs are stored with the outputs so
the fixtures are self-contained.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402
from oracle import np_oracle as O    # noqa: E402  (only for the synthetic input generator)

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **kw):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **kw)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    H.install()
    from DistantSpeech.transform.transform import Transform, stft, istft
    from DistantSpeech.beamformer.MicArray import MicArray
    from DistantSpeech.beamformer.beamformer import beamformer, compute_mvdr_weight
    from DistantSpeech.beamformer.fixedbeamformer import FixedBeamformer
    from DistantSpeech.noise_estimation.mcspp_base import McSppBase
    from DistantSpeech.noise_estimation.mcra import NoiseEstimationMCRA

    rng = np.random.default_rng(1234)

    # ---- a1/a2: module-level stft / istft ------------------------------------
    x = (rng.standard_normal(6000) * 0.1).astype(np.float32)
    win = O.sqrt_hann(512)
    D = stft(x.astype(np.float64), n_fft=512, hop_length=128, window=win, center=True)
    y = istft(D, hop_length=128, window=win, center=True, length=6000)
    D2 = stft(x.astype(np.float64), n_fft=256, hop_length=128, window=O.sqrt_hann(256), center=False)
    y2 = istft(D2, hop_length=128, window=O.sqrt_hann(256), center=False)
    save("stft_istft.npz", x=x, D_512_128_center=D, y_512_128_center=y, D_256_128_plain=D2, y_256_128_plain=y2)

    # ---- a3: streaming Transform, 3 chunks -----------------------------------
    xs = (rng.standard_normal((256 * 12, 3)) * 0.1).astype(np.float32)
    tf = Transform(n_fft=512, hop_length=256, channel=3)
    chunks = [xs[:256 * 5], xs[256 * 5:256 * 6], xs[256 * 6:]]
    Ys = [tf.stft(c.astype(np.float64)) for c in chunks]
    tf2 = Transform(n_fft=512, hop_length=256, channel=3)
    ys = [np.atleast_2d(tf2.istft(Y)) for Y in Ys]
    save("transform_stream.npz", x=xs, Y0=Ys[0], Y1=Ys[1], Y2=Ys[2], y0=ys[0], y1=ys[1], y2=ys[2],
         prev_in=tf.previous_input, prev_out=tf2.previous_output)

    # ---- a9: fixed beamformer -------------------------------------------------
    geo8 = O.MicGeometry("circular", r=0.05, M=8, n_fft=256)
    x8 = O.synth_streams(1, geo8, 8192, seed0=77)[0].T.copy()            # [N, 8] float32
    mic8 = MicArray(arrayType="circular", r=0.05, M=8, n_fft=256)
    fb = FixedBeamformer(mic8, 256, 128, 256)
    y_sd = fb.process(x8.astype(np.float64), (30, 0))                     # reference always uses SD here
    W_sd = fb.W.copy()
    fb2 = FixedBeamformer(mic8, 256, 128, 256)
    W_ds = beamformer.compute_weights(fb2, (30, 0), "DS")
    D8 = fb2.transform.stft(x8.astype(np.float64))
    Yf = np.einsum("km,ktm->kt", W_ds.conj(), D8)[:, :, None]
    y_ds = fb2.transform.istft(Yf)
    save("fixedbf.npz", x=x8, angle=np.array([30, 0]), W_sd=W_sd, y_sd=y_sd, W_ds=W_ds, y_ds=y_ds)

    # ---- a12: MCRA -------------------------------------------------------------
    P = np.abs(D8[:, :, 0]) ** 2                                           # [129, T]
    m = NoiseEstimationMCRA(nfft=256)
    lam = np.zeros_like(P)
    pp = np.zeros_like(P)
    for n in range(P.shape[1]):
        lam[:, n] = m.estimation(P[:, n])
        pp[:, n] = m.p
    save("mcra.npz", P=P, lambda_d=lam, p=pp, S=m.S, Smin=m.Smin, Stmp=m.Stmp, ell=m.ell, frm_cnt=m.frm_cnt)

    # ---- a13 + config-4 composition ---------------------------------------------
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xc = O.synth_streams(1, geo, 16384, seed0=99)[0].T.copy()             # [N, 8]
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    T = Transform(n_fft=512, hop_length=256, channel=8)
    Dc = T.stft(xc.astype(np.float64))
    est = McSppBase(nfft=512, channels=8)
    a0 = beamformer(mic, frame_len=512, hop=256, nfft=512).compute_steering_vector_from_doa((30, 0))
    nT = Dc.shape[1]
    Y = np.zeros((257, nT, 1), dtype=complex)
    tr = {k: np.zeros((257, nT)) for k in ("p", "xi", "gamma", "q", "G")}
    for n in range(nT):
        yv = Dc[:, n, :]
        est.estimation(yv)
        w = compute_mvdr_weight(a0, est.Phi_vv_inv)
        est.compute_omlsa_weight(est.xi, est.p)
        Y[:, n, 0] = np.einsum("ij,ij->i", w.conj(), yv) * est.G
        for k in tr:
            tr[k][:, n] = getattr(est, k)
    yc = Transform(n_fft=512, hop_length=256, channel=1).istft(Y)
    save("chain_mcspp_mvdr.npz", x=xc, look=np.array([30, 0]), a0=a0, y=yc, Yspec=Y[:, :, 0].astype(np.complex64),
         p=tr["p"].astype(np.float32), xi=tr["xi"], gamma=tr["gamma"], q=tr["q"].astype(np.float32),
         G=tr["G"].astype(np.float32), Phi_vv_last=est.Phi_vv, Phi_yy_last=est.Phi_yy, w_pmwf_last=est.w,
         Phi_vv_inv_last=est.Phi_vv_inv)

    # ---- a11: online MVDR (run_MVDRbeamformer.py path) ----------------------------
    geo4 = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    x4 = O.synth_streams(1, geo4, 12800, seed0=55)[0].copy()               # [4, N]
    mic4 = MicArray(arrayType="circular", r=0.032, M=4, n_fft=256)
    ab = H.make_adaptive_mvdr(mic4, 256, 128, 256)
    ang = np.array([30, 0]) / 180 * np.pi
    with np.errstate(all="ignore"):
        ya = ab.process(x4.astype(np.float64), ang, method=2)["data"]
    save("adaptive_mvdr.npz", x=x4, angle_rad=ang, y=ya, H_last=ab.H, Rvv_last=ab.Rvv, p_last=ab.mcra.p)

    # ---- a17: FDGSC (config 3 shape: 6-mic linear r=0.05, frameLen 256) ------------
    import contextlib
    import io
    from DistantSpeech.beamformer.FDGSC import FDGSC
    geo6 = O.MicGeometry("linear", r=0.05, M=6, n_fft=256)
    x6 = O.synth_streams(1, geo6, 256 * 60, look_deg=(60.0, 0.0), interf_deg=(140.0, 0.0), seed0=33)[0].T.copy()
    mic6 = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints tau / filter shapes
        fd = FDGSC(mic6, frameLen=256, angle=[60, 0])
    xin = x6.astype(np.float64)
    res = fd.process(xin, postfilter=False, dc_notch=True)   # mutates xin (DC notch in place)
    save("fdgsc.npz", x=x6, angle_deg=np.array([60, 0]), y=res[0], p=res[1].astype(np.float32),
         fix_output=res[2].astype(np.float32), bm_output=res[4].astype(np.float32),
         x_notched=xin.astype(np.float32), delay_filter=fd.time_alignment.delay_filter,
         W_aic_last=fd.aic_filter.W, W_bm0_last=fd.bm[0].W)

    # ---- a17 with postfilter=True (two calls: the second sees the carried transform / OMLSA state) ----
    with contextlib.redirect_stdout(io.StringIO()):
        fdp = FDGSC(mic6, frameLen=256, angle=[60, 0])
    xa, xb = x6[:256 * 40].astype(np.float64), x6[256 * 40:].astype(np.float64)
    ya = fdp.process(xa, postfilter=True, dc_notch=True)[0]
    yb = fdp.process(xb, postfilter=True, dc_notch=True)[0]
    save("fdgsc_postfilter.npz", n_first=np.array(256 * 40), n_total=np.array(256 * 60), y=np.concatenate([ya, yb]),
         G_last=fdp.omlsa_multi.G)

    # ---- a15: NsOmlsaMulti on the FDGSC outputs, a16: Zelinski postfilter weights -------
    from DistantSpeech.noise_estimation.omlsa_multi import NsOmlsaMulti
    from DistantSpeech.postfilter.postfilter import PostFilter
    tfy = Transform(n_fft=512, hop_length=256, channel=1)
    tfu = Transform(n_fft=512, hop_length=256, channel=5)
    Yp = (np.abs(tfy.stft(res[0])[:, :, 0]) ** 2).astype(np.float32).astype(np.float64)      # [257, T]
    Up = (np.abs(tfu.stft(res[4][:, :5])) ** 2).astype(np.float32).astype(np.float64)        # [257, T, 5]
    om = NsOmlsaMulti(nfft=512, cal_weights=True, M=6)
    Gs, lams, ps = [], [], []
    for n in range(Yp.shape[1]):
        om.estimation(Yp[:, n], Up[:, n, :])
        Gs.append(om.G.copy()); lams.append(np.array(om.lambda_d, dtype=float).copy()); ps.append(om.p.copy())
    save("omlsa_multi.npz", Y=Yp.astype(np.float32), U=Up.astype(np.float32), G=np.array(Gs), lambda_d=np.array(lams),
         p=np.array(ps))
    pf = PostFilter(mic8, 256, 128, 256)
    Z8 = D8[:, :40, :].transpose(2, 0, 1)                              # [M, K, T] from the fixed-BF fixture
    Ws = np.array([pf.getweights(Z8[:, :, n]).squeeze() for n in range(40)])
    save("zelinski.npz", Z=Z8.astype(np.complex64), W=Ws, Pxii=pf.Pxii, Pxij=pf.Pxij)

    # ---- f1: frequency-domain GSC with the McMcra postfilter (GSC.py:174-294), 4 mics, two calls ----
    from DistantSpeech.beamformer.MicArray import MicArray as RefMic
    geo_g = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    xg = O.synth_streams(1, geo_g, 128 * 160, seed0=0x65C)[0]                                   # [4, N] float32
    gsc = H.make_gsc(RefMic(arrayType="circular", r=0.032, M=4), 256)
    ang = np.array([30, 0]) / 180 * np.pi
    n1 = 128 * 100
    ps, Gs, ys = [], [], []
    for lo, hi in ((0, n1), (n1, xg.shape[1])):
        ys.append(gsc.process(xg[:, lo:hi].astype(np.float64), ang, method=2)["data"])
    save("gsc.npz", x=xg, angle_rad=ang, n_first=np.array(n1), y=np.concatenate(ys), p_last=gsc.spp.p, q_last=gsc.spp.q,
         xi_last=gsc.spp.xi, gamma_last=gsc.spp.gamma, Gpost_last=gsc.spp.G, Gw_last=gsc.G, Phi_yy_last=gsc.spp.Phi_yy,
         Phi_vv_last=gsc.spp.Phi_vv)

    # ---- f1: SubbandGSC (STFT-domain NLMS blocking filters + canceller, McSpp gate), 4 mics, two calls ----
    xs4 = O.synth_streams(1, geo_g, 256 * 70 + 77, seed0=0x5B6)[0]                              # [4, N] float32
    sgsc = H.make_subband_gsc(RefMic(arrayType="circular", r=0.032, M=4), 256, angle=[30, 0])
    n1s = 256 * 40 + 77
    outs = []
    for lo, hi in ((0, n1s), (n1s, xs4.shape[1])):
        with contextlib.redirect_stdout(io.StringIO()):
            outs.append(sgsc.process(xs4[:, lo:hi].astype(np.float64)))
    save("subband_gsc.npz", x=xs4, n_first=np.array(n1s), y=np.concatenate([o[0] for o in outs]),
         fix_output=np.concatenate([o[1] for o in outs]).astype(np.float32),
         bm_output=np.concatenate([o[2] for o in outs]).astype(np.float32),
         p=np.concatenate([o[3] for o in outs], axis=1).astype(np.float32),
         W_aic_last=sgsc.aic_filter.W, W_bm0_last=sgsc.bm[0].W)

    # ---- f1: SubbandRLS, 2 taps, 14 blocks -------------------------------------------------------------------
    from DistantSpeech.adaptivefilter.SubbandRLS import SubbandRLS
    rng = np.random.default_rng(0x715)
    xr = (rng.standard_normal(256 * 14) * 0.2).astype(np.float32)
    dr = (0.5 * np.roll(xr, 5) + 0.05 * rng.standard_normal(256 * 14)).astype(np.float32)
    rls = SubbandRLS(filter_len=2, num_bands=512)
    errs = [rls.update(xr[256 * n:256 * (n + 1)].astype(np.float64), dr[256 * n:256 * (n + 1)].astype(np.float64))[0]
            for n in range(14)]
    save("subband_rls.npz", x=xr, d=dr, err=np.concatenate(errs), W_last=rls.W, P_last=rls.P)

    # ---- f1: TDGSC (fixed blocking matrix + constrained FDAF canceller), 4 mics, two calls with a ragged tail ----
    from DistantSpeech.beamformer.TDGSC import TDGSC
    xt4 = np.ascontiguousarray(O.synth_streams(1, geo_g, 256 * 70 + 50, seed0=0x7D6)[0].T)       # [N, 4] float32
    with contextlib.redirect_stdout(io.StringIO()):
        td = TDGSC(RefMic(arrayType="circular", r=0.032, M=4), frameLen=256, angle=[30, 0])
    n1t = 256 * 40 + 50
    touts = []
    for lo, hi in ((0, n1t), (n1t, xt4.shape[0])):
        with contextlib.redirect_stdout(io.StringIO()):
            touts.append(td.process(xt4[lo:hi].astype(np.float64)))
    save("tdgsc.npz", x=xt4, n_first=np.array(n1t), y=np.concatenate([o[0] for o in touts]),
         p=np.concatenate([o[1] for o in touts], axis=1).astype(np.float32),
         bm_output=np.concatenate([o[2] for o in touts]).astype(np.float32), W_last=td.aic_filter.W)

    # ---- a14: McSpp (CDR-driven prior, complex inverse with SNR-dependent loading), 4 mics ----
    from DistantSpeech.noise_estimation.mcspp import McSpp
    geo4 = O.MicGeometry("circular", r=0.032, M=4, n_fft=512)
    x4 = np.ascontiguousarray(O.synth_streams(1, geo4, 256 * 240, seed0=0xCD4)[0].T)           # [N, 4] float32
    D4 = Transform(n_fft=512, hop_length=256, channel=4).stft(x4.astype(np.float64))          # [257, 240, 4]
    with contextlib.redirect_stdout(io.StringIO()):
        est = McSpp(nfft=512, channels=4)
    taps = {k: [] for k in ("p", "xi", "gamma", "q")}
    for n in range(D4.shape[1]):
        est.estimation(D4[:, n, :])
        taps["p"].append(est.p.copy()); taps["xi"].append(est.xi.copy()); taps["gamma"].append(est.gamma.copy())
        taps["q"].append(est.q.copy())
    save("mcspp_cdr.npz", x=x4, **{k: np.array(v).T for k, v in taps.items()}, w_last=est.w, Phi_yy_last=est.Phi_yy,
         Phi_vv_last=est.Phi_vv, Phi_vv_inv_last=est.Phi_vv_inv, Phi_xx_last=est.Phi_xx,
         mcra_p_last=est.mccdr.mcra.p, Pxii_last=est.mccdr.Gamma_estimator.Pxii, Pxij_last=est.mccdr.Gamma_estimator.Pxij)

    mask_beamformers()
    idoa()


def mask_beamformers():
    """8f.3: mask-based MVDR (mvdr.ipynb cell 6) and GEV (cell 8) on a 6-mic array like the notebook's, with the
    mask from the reference's McSppBase (the notebook's McSpp raises above 4 microphones)."""
    H.install()
    from DistantSpeech.transform.transform import Transform
    from DistantSpeech.beamformer.beamformer import (steering, compute_mvdr_weight, get_gev_vector, phase_correction,
                                                     blind_analytic_normalization)
    from DistantSpeech.noise_estimation.mcspp_base import McSppBase
    geo6 = O.MicGeometry("circular", r=0.05, M=6, n_fft=512)
    x6 = np.ascontiguousarray(O.synth_streams(1, geo6, 256 * 100, seed0=0x6E7)[0].T)           # [N, 6] float32
    tf = Transform(n_fft=512, hop_length=256, channel=6)
    D = tf.stft(x6.astype(np.float64))                                                         # [257, 100, 6]
    K, nT, M = D.shape
    est = McSppBase(nfft=512, channels=6)
    p = np.zeros((K, nT))
    Pxx = np.zeros((K, M, M), dtype=complex)
    Pvv = np.zeros((K, M, M), dtype=complex)
    for n in range(nT):
        est.estimation(D[:, n, :])
        p[:, n] = est.p
    for n in range(nT):                                                                       # cell 6
        y = D[:, n, :]
        Pxx = Pxx + np.einsum('ij,il->ijl', y, y.conj()) * p[:, n:n + 1, None]
        Pvv = Pvv + np.einsum('ij,il->ijl', y, y.conj()) * (1 - p[:, n:n + 1, None])
    steer = steering(Pxx)
    w_mvdr = compute_mvdr_weight(steer, np.linalg.inv(Pvv))
    y_mvdr = Transform(n_fft=512, hop_length=256, channel=1).istft(np.einsum('inj,ij->in', D, w_mvdr.conj())[:, :, None])
    w_gev_raw = get_gev_vector(Pxx, Pvv)                                                       # cell 8
    w_gev_pc = phase_correction(w_gev_raw)
    w_gev = blind_analytic_normalization(w_gev_pc, Pvv)
    y_gev = Transform(n_fft=512, hop_length=256, channel=1).istft(np.einsum('inj,ij->in', D, w_gev.conj())[:, :, None])
    # the frame-range averages + GEVD-flavoured steering of cell 2 (steering() on a non-Hermitian product)
    Rvv = sum(np.einsum('ij,il->ijl', D[:, n, :], D[:, n, :].conj()) for n in range(10)) / 10
    Ryy = sum(np.einsum('ij,il->ijl', D[:, n, :], D[:, n, :].conj()) for n in range(30, 80)) / 50
    steer_pca = steering(Ryy - Rvv)
    save("mask_beamformers.npz", x=x6, p=p, Pxx=Pxx, Pvv=Pvv, steer=steer, w_mvdr=w_mvdr,
         y_mvdr=y_mvdr, w_gev_raw=w_gev_raw, w_gev_pc=w_gev_pc, w_gev=w_gev, y_gev=y_gev, Rvv=Rvv, Ryy=Ryy,
         steer_pca=steer_pca)


def idoa():
    """8f.4: Idoa.estimate / Idoa.process (doa/idoa.py) on a 4-mic circular and a 6-mic linear array."""
    import contextlib
    import io
    H.install()
    from DistantSpeech.doa.idoa import Idoa
    from DistantSpeech.beamformer.MicArray import MicArray
    out = {}
    for tag, arr, M, r, n_fft in (("c4", "circular", 4, 0.032, 256), ("l6", "linear", 6, 0.05, 512)):
        geo = O.MicGeometry(arr, r=r, M=M, n_fft=n_fft)
        x = np.ascontiguousarray(O.synth_streams(1, geo, (n_fft // 2) * 60, seed0=0x1D0A + M)[0].T)   # [N, M] float32
        with contextlib.redirect_stdout(io.StringIO()):
            mic = MicArray(arrayType=arr, r=r, M=M, n_fft=n_fft)
            a, b, c = Idoa(mic), Idoa(mic), Idoa(mic)
        n1 = (n_fft // 2) * 25
        y = np.concatenate([a.process(x[:n1].astype(np.float64), default_direction=30),
                            a.process(x[n1:].astype(np.float64), default_direction=30)])          # two chunks: state carries over
        X = b.transform.stft(x.astype(np.float64))
        p_all = b.estimate(X)
        p_one = c.estimate(X, theta=40)
        sel = np.array([30, 40, 41, 179])
        out.update({tag + "_x": x, tag + "_n1": np.array(n1), tag + "_y": y, tag + "_sel": sel, tag + "_p_sel": p_all[:, :, sel],
                    tag + "_p_theta40": p_one[:, :, [40, 41]], tag + "_mu_Delta_last": b.mu_Delta[:, sel],
                    tag + "_var_last": b.var_Delta_h0[:, sel], tag + "_Psi_sel": b.Psi[:, :, sel]})
    save("idoa.npz", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mask":
        mask_beamformers()
    elif len(sys.argv) > 1 and sys.argv[1] == "idoa":
        idoa()
    else:
        main()

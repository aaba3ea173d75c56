"""numpy emulator of the device tap-GEMM primitive kp_tapconv_bf16 (TEST INFRASTRUCTURE ONLY).

Executes a TapPlan exactly as include/kp_b200.h specifies it (strided views with zero fill outside,
per-tap shifts, packed K-major weights, strided output view) so that the host-side lowering in
kp_b200/tapconv.py can be verified on a CPU box against the convolution oracle (oracle/tf_ops.py).
"""
import numpy as np


def _view_read(flat, v, N, n, h, w):
    """Read view element block [C] at (n,h,w); zeros outside the view extent."""
    if not (0 <= n < N and 0 <= h < v["Hd"] and 0 <= w < v["Wd"]):
        return np.zeros(v["C"], dtype=flat.dtype)
    base = v["off"] + n * v["sn"] + h * v["sh"] + w * v["sw"]
    return flat[base:base + v["C"]]


def run_plan(plan, srcs, wpacked, bias, out, act=None):
    """srcs: list of numpy NHWC arrays; wpacked [rows_pad,Ktot]; out: numpy array written in place through the
    plan's output view (flattened)."""
    flats = [np.ascontiguousarray(s).reshape(-1) for s in srcs]
    oflat = out.reshape(-1)
    CB = plan.CB
    N = plan.N
    # vectorised over the output grid: gather A as [N,Ho,Wo,Ktot]
    Ho, Wo = plan.Ho, plan.Wo
    A = np.zeros((N, Ho, Wo, plan.Ktot), dtype=wpacked.dtype)
    col = 0
    for (dh, dw, mf) in plan.taps:
        for s in range(plan.n_src):
            v = plan.views[mf + s]
            flat = flats[v["src"]]
            C = v["C"]
            hh = np.arange(Ho) + dh
            ww = np.arange(Wo) + dw
            okh = (hh >= 0) & (hh < v["Hd"])
            okw = (ww >= 0) & (ww < v["Wd"])
            n_idx = np.arange(N)
            base = (v["off"] + n_idx[:, None, None] * v["sn"] + np.clip(hh, 0, None)[None, :, None] * v["sh"]
                    + np.clip(ww, 0, None)[None, None, :] * v["sw"])
            ok = okh[None, :, None] & okw[None, None, :] & np.ones((N, 1, 1), dtype=bool)
            base = np.where(ok, base, 0)
            vals = flat[base[..., None] + np.arange(C)[None, None, None, :]]
            vals = np.where(ok[..., None], vals, 0)
            A[..., col:col + C] = vals
            col += -(-C // CB) * CB
    assert col == plan.Ktot
    Y = A.reshape(-1, plan.Ktot) @ wpacked[:plan.rows].T
    if bias is not None:
        Y = Y + bias[:plan.rows]
    if act is not None:
        Y = act(Y)
    Y = Y.reshape(N, Ho, Wo, plan.rows)
    n_idx, h_idx, w_idx = np.meshgrid(np.arange(N), np.arange(Ho), np.arange(Wo), indexing="ij")
    obase = plan.out_off + n_idx * plan.out_sn + h_idx * plan.out_sh + w_idx * plan.out_sw
    oflat[obase[..., None] + np.arange(plan.rows)[None, None, None, :]] = Y
    return out


def run_wgrad_plan(plan, x, dy, dw):
    """Emulates kp_tapconv_wgrad_bf16: dw (flattened HWIO, f64) += sum_pixels X_tap^T dY."""
    xf = np.ascontiguousarray(x).reshape(-1)
    dyf = np.ascontiguousarray(dy).reshape(-1)
    dwf = dw.reshape(-1)
    N, Ho, Wo = plan.N, plan.Ho, plan.Wo

    def gather(flat, v, dh, dw_):
        C = v["C"]
        hh = np.arange(Ho) + dh
        ww = np.arange(Wo) + dw_
        ok = ((hh >= 0) & (hh < v["Hd"]))[None, :, None] & ((ww >= 0) & (ww < v["Wd"]))[None, None, :] \
            & np.ones((N, 1, 1), dtype=bool)
        base = (v["off"] + np.arange(N)[:, None, None] * v["sn"] + np.clip(hh, 0, None)[None, :, None] * v["sh"]
                + np.clip(ww, 0, None)[None, None, :] * v["sw"])
        base = np.where(ok, base, 0)
        vals = flat[base[..., None] + np.arange(C)[None, None, None, :]]
        return np.where(ok[..., None], vals, 0).reshape(-1, C)

    DY = gather(dyf, plan.dy, 0, 0)[:, :plan.Cout]
    for (dh, dw_, mf, tf) in plan.taps:
        X = gather(xf, plan.views[mf], dh, dw_)[:, :plan.Cin]
        G = X.T @ DY
        idx = plan.dw_off + tf * plan.dw_stap + np.arange(plan.Cin)[:, None] * plan.dw_sci + np.arange(plan.Cout)[None, :]
        dwf[idx] += G
    return dw

"""Bring-up probe for the CUDA kernels: each case runs in its own subprocess (a trapped kernel
poisons the CUDA context) under a timeout and reports one JSON object; the parent writes
gpurun_out/probe.json.  Run on the GPU box:

    python tools/gpu_probe.py [case ...]

This is a development aid (diagnostics are more verbose than the pytest parity tests).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _err_report(got, ref, name, tol):
    import torch
    got = got.float()
    ref = ref.float()
    d = (got - ref).abs()
    out = {"case": name, "max_abs": float(d.max()), "mean_abs": float(d.mean()), "ref_absmax": float(ref.abs().max()),
           "nan": int(torch.isnan(got).sum()), "ok": bool(d.max() <= tol and not torch.isnan(got).any())}
    if not out["ok"]:
        bad = d > tol
        out["bad_frac"] = float(bad.float().mean())
        if got.dim() == 2:
            M, N = got.shape
            rb = bad.float().reshape(-1, N).mean(1)
            cb = bad.float().mean(0)
            out["bad_rows_first64"] = [round(float(x), 2) for x in rb[:64]]
            out["bad_cols_first128"] = [round(float(x), 2) for x in cb[:128]]
            out["got_00"] = [[round(float(v), 4) for v in got[i, :8]] for i in range(min(4, M))]
            out["ref_00"] = [[round(float(v), 4) for v in ref[i, :8]] for i in range(min(4, M))]
    return out


def case_elementwise():
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    res = []
    M = 1000
    x = torch.randn(M, 768, device=dev).bfloat16()
    g = (1 + 0.1 * torch.randn(768, device=dev)).float()
    b = (0.1 * torch.randn(768, device=dev)).float()
    y, mean, rstd = ops.layernorm(x, g, b, save_stats=True)
    ref = torch.nn.functional.layer_norm(x.float(), (768,), g, b, 1e-12)
    res.append(_err_report(y, ref, "layernorm_fwd", 2e-2))
    # LN bwd
    dy = torch.randn(M, 768, device=dev).bfloat16()
    dg = torch.zeros(768, device=dev)
    db = torch.zeros(768, device=dev)
    dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db)
    xr = x.float().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (768,), gr, br, 1e-12).backward(dy.float())
    res.append(_err_report(dx, xr.grad, "layernorm_bwd_dx", 3e-2))
    res.append(_err_report(dg, gr.grad, "layernorm_bwd_dgamma", 2e-2 * gr.grad.abs().max().item()))
    res.append(_err_report(db, br.grad, "layernorm_bwd_dbeta", 1e-3 * br.grad.abs().max().item() + 1e-3))
    # embed text
    B, S, V = 3, 256, 5000
    word = torch.randn(V, 768, device=dev)
    pos = torch.randn(512, 768, device=dev)
    typ = torch.randn(2, 768, device=dev)
    ids_full = torch.randint(0, V, (B, 512), device=dev)
    out = ops.embed_text_ln(ids_full[:, :256], word, pos, typ, g, b)
    ref = torch.nn.functional.layer_norm((word[ids_full[:, :256]] + typ[0]) + pos[:256], (768,), g, b, 1e-12)
    res.append(_err_report(out.view(B, S, 768), ref, "embed_text_ln", 3e-2))
    # embed joint
    N = 3000
    table = torch.randn(N + 3, 768, device=dev)
    lm = torch.randn(B, 256, 768, device=dev).bfloat16()
    ids = torch.randint(0, N + 3, (B, 512), device=dev)
    tt = torch.cat([torch.zeros(B, 256, dtype=torch.long), torch.ones(B, 256, dtype=torch.long)], 1).to(dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    out, mean, rstd, emb = ops.embed_joint_ln(ids, tt, lm, table, pos, typ, g, b, save_stats=True,
                                              want_inputs_embeds=True, err_flag=flag)
    src = torch.cat([lm.float(), table[ids[:, 256:]]], 1)
    res.append({"case": "embed_joint_gather_bitexact", "ok": bool(torch.equal(emb.view(B, 512, 768), src)),
                "flag": int(flag.item())})
    ref = torch.nn.functional.layer_norm((src + typ[tt]) + pos, (768,), g, b, 1e-12)
    res.append(_err_report(out.view(B, 512, 768), ref, "embed_joint_ln", 3e-2))
    out2, _, _, _ = ops.embed_joint_ln(ids, None, lm, table, pos, typ, g, b)
    res.append({"case": "embed_joint_default_types", "ok": bool(torch.equal(out, out2))})
    bad = ids.clone()
    bad[0, 300] = N + 3
    ops.embed_joint_ln(bad, tt, lm, table, pos, typ, g, b, err_flag=flag)
    res.append({"case": "embed_joint_oob_flag", "ok": int(flag.item()) == 1})
    # embed joint bwd
    dy = torch.randn(B * 512, 768, device=dev).bfloat16()
    dpos = torch.zeros(512, 768, device=dev)
    dtyp = torch.zeros(2, 768, device=dev)
    dg = torch.zeros(768, device=dev)
    db = torch.zeros(768, device=dev)
    ops.embed_joint_ln_bwd(ids, tt, lm, table, pos, typ, g, mean, rstd, dy, dpos, dtyp, dg, db)
    pr = pos.clone().requires_grad_(True)
    tr = typ.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm((src + tr[tt]) + pr, (768,), gr, br, 1e-12).backward(dy.float().view(B, 512, 768))
    res.append(_err_report(dpos, pr.grad, "embed_joint_bwd_dpos", 1e-3 * pr.grad.abs().max().item() + 1e-4))
    res.append(_err_report(dtyp, tr.grad, "embed_joint_bwd_dtype", 1e-3 * tr.grad.abs().max().item() + 1e-3))
    res.append(_err_report(dg, gr.grad, "embed_joint_bwd_dgamma", 1e-3 * gr.grad.abs().max().item() + 1e-3))
    res.append(_err_report(db, br.grad, "embed_joint_bwd_dbeta", 1e-3 * br.grad.abs().max().item() + 1e-3))
    # misc
    w = torch.randn(1000, 77, device=dev)
    res.append({"case": "cast_bf16", "ok": bool(torch.equal(ops.cast_bf16(w), w.bfloat16()))})
    m = torch.randint(0, 2, (4, 512), device=dev)
    ref = (1.0 - m.float()) * torch.finfo(torch.float32).min
    res.append({"case": "mask_to_bias", "ok": bool(torch.equal(ops.mask_to_bias(m), ref))})
    idx = torch.randperm(M, device=dev)[:300].int()
    res.append({"case": "gather_rows", "ok": bool(torch.equal(ops.gather_rows(x, idx), x[idx.long()]))})
    dst = torch.randn(M, 768, device=dev).bfloat16()
    want = dst.clone()
    want[idx.long()] = (want[idx.long()].float() + x[:300].float()).bfloat16()
    ops.scatter_add_rows(x[:300].contiguous(), idx, dst)
    res.append({"case": "scatter_add_rows", "ok": bool(torch.equal(dst, want))})
    cs = torch.empty(768, device=dev)
    ops.colsum(x, cs)
    res.append(_err_report(cs, x.float().sum(0), "colsum", 1e-2))
    pooled = torch.randn(5, 768, device=dev)
    wn = torch.randn(2, 768, device=dev) * 0.05
    bn = torch.randn(2, device=dev)
    lab = torch.randint(0, 2, (5,), device=dev)
    lg, rl = ops.nsp_head(pooled, wn, bn, lab)
    lref = pooled @ wn.T + bn
    res.append(_err_report(lg, lref, "nsp_logits", 1e-4))
    res.append(_err_report(rl, torch.nn.functional.cross_entropy(lref, lab, reduction="none"), "nsp_row_loss", 1e-4))
    return res


def _gemm_ref(a, b, a_major, b_major):
    A = a.float() if a_major == 0 else a.float().T
    Bm = b.float() if b_major == 0 else b.float().T
    return A @ Bm.T


def _mk(rows, cols, dev, scale=1.0):
    import torch
    return (torch.randn(rows, cols, device=dev) * scale).bfloat16()


def case_gemm_basic():
    """Smallest possible: one tile, K = 64 (one k-block), then a few more shapes; fp32 output."""
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(1)
    res = []
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 768), (384, 768, 3072), (1000, 1000, 520)]:
        a = _mk(M, K, "cuda")
        b = _mk(N, K, "cuda")
        c = ops.gemm(a, b, M=M, N=N, K=K, epilogue=ops.EPI_F32)
        torch.cuda.synchronize()
        res.append(_err_report(c, _gemm_ref(a, b, 0, 0), f"gemm_kk_f32_{M}x{N}x{K}", 1e-2 * (K ** 0.5) / 8))
    return res


def case_gemm_epilogues():
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(2)
    res = []
    M, N, K = 520, 768, 768
    a = _mk(M, K, "cuda", 0.5)
    w = _mk(N, K, "cuda", 0.05)
    bias = torch.randn(N, device="cuda") * 0.5
    acc = _gemm_ref(a, w, 0, 0)
    tol = 3e-2
    res.append(_err_report(ops.linear(a, w, bias), acc + bias, "epi_bias", tol))
    res.append(_err_report(ops.linear(a, w, None), acc, "epi_nobias", tol))
    res.append(_err_report(ops.linear(a, w, bias, ops.EPI_BIAS_GELU), torch.nn.functional.gelu(acc + bias),
                           "epi_bias_gelu", tol))
    pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    act = ops.linear(a, w, bias, ops.EPI_BIAS_GELU_SAVE, c2=pre)
    res.append(_err_report(act, torch.nn.functional.gelu(acc + bias), "epi_gelu_save_act", tol))
    res.append(_err_report(pre, acc + bias, "epi_gelu_save_pre", tol))
    r = _mk(M, N, "cuda")
    res.append(_err_report(ops.linear(a, w, bias, ops.EPI_BIAS_RESID, resid=r), acc + bias + r.float(),
                           "epi_bias_resid", 4e-2))
    res.append(_err_report(ops.linear(a, w, bias, ops.EPI_BIAS_TANH_F32), torch.tanh(acc + bias), "epi_tanh_f32", 2e-3))
    u = _mk(M, N, "cuda")
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf).sum().backward()
    wT = w.T.contiguous()  # stored [K][N]: the dgrad layout
    res.append(_err_report(ops.gemm(a, wT, M=M, N=N, K=K, b_major=1, epilogue=ops.EPI_DGELU, resid=u), acc * uf.grad,
                           "epi_dgelu", tol))
    # training forward of BertIntermediate: activation + SAVED DERIVATIVE, and the backward that multiplies by it
    dsave = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    act2 = ops.linear(a, w, bias, ops.EPI_BIAS_GELU_SAVE_GRAD, c2=dsave)
    pre_f = (acc + bias).detach().requires_grad_(True)
    torch.nn.functional.gelu(pre_f).sum().backward()
    res.append(_err_report(act2, torch.nn.functional.gelu(acc + bias), "epi_gelu_save_grad_act", tol))
    res.append(_err_report(dsave, pre_f.grad, "epi_gelu_save_grad_derivative", 1e-2))
    res.append(_err_report(ops.gemm(a, wT, M=M, N=N, K=K, b_major=1, epilogue=ops.EPI_MUL, resid=u), acc * u.float(), "epi_mul", tol))
    # activation and derivative over the whole useful range: acc[m, n] = x_m exactly (bf16 grid), bias 0
    xs = torch.linspace(-9.0, 9.0, 4096, device="cuda").bfloat16()
    a_x = torch.zeros(4096, 64, device="cuda").bfloat16()
    a_x[:, 0] = xs
    w_1 = torch.zeros(64, 64, device="cuda").bfloat16()
    w_1[:, 0] = 1.0
    dgrid = torch.empty((4096, 64), dtype=torch.bfloat16, device="cuda")
    ygrid = ops.linear(a_x, w_1, torch.zeros(64, device="cuda"), ops.EPI_BIAS_GELU_SAVE_GRAD, c2=dgrid)
    xg = xs.float().requires_grad_(True)
    yref = torch.nn.functional.gelu(xg)
    yref.sum().backward()
    res.append(_err_report(ygrid[:, 5], yref.detach(), "gelu_grid_activation", 4e-2))
    res.append(_err_report(dgrid[:, 5], xg.grad, "gelu_grid_derivative", 6e-3))
    # strided A (pooler reads row 0 of every sequence): lda = 512*768
    seq = _mk(4 * 512, 768, "cuda", 0.5)
    a0 = seq.view(4, 512, 768)[:, 0]
    res.append(_err_report(ops.gemm(a0, w, M=4, N=N, K=K, epilogue=ops.EPI_BIAS_TANH_F32, bias=bias),
                           torch.tanh(a0.float() @ w.float().T + bias), "epi_tanh_strided_rows", 2e-3))
    return res


def case_gemm_ln():
    """Fused bias + residual + LayerNorm epilogue (3-CTA clusters, DSMEM row statistics) vs torch fp32."""
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(7)
    res = []
    for (M, K) in [(128, 768), (520, 768), (1000, 3072), (128 * 160 + 37, 768)]:
        N = 768
        a = _mk(M, K, "cuda", 0.5)
        w = _mk(N, K, "cuda", 0.05)
        bias = torch.randn(N, device="cuda") * 0.5
        r = _mk(M, N, "cuda")
        gamma = 1.0 + 0.2 * torch.randn(N, device="cuda")
        beta = 0.3 * torch.randn(N, device="cuda")
        zref = _gemm_ref(a, w, 0, 0) + bias + r.float()
        yref = torch.nn.functional.layer_norm(zref, (N,), gamma, beta, 1e-12)
        y = ops.linear_resid_ln(a, w, bias, r, gamma, beta)
        torch.cuda.synchronize()
        res.append(_err_report(y, yref, f"ln_infer_{M}x{K}", 6e-2))
        y2, z, mean, rstd = ops.linear_resid_ln(a, w, bias, r, gamma, beta, save_for_backward=True)
        torch.cuda.synchronize()
        res.append(_err_report(y2, yref, f"ln_train_y_{M}x{K}", 6e-2))
        res.append(_err_report(z, zref, f"ln_train_z_{M}x{K}", 5e-2))
        res.append(_err_report(mean, zref.mean(1), f"ln_train_mean_{M}x{K}", 2e-3))
        res.append(_err_report(rstd, (zref.var(1, unbiased=False) + 1e-12).rsqrt(), f"ln_train_rstd_{M}x{K}", 2e-3))
        res.append({"case": f"ln_same_{M}x{K}", "ok": bool(torch.equal(y, y2))})
    # A and the residual read through a row pitch (row b*512 of a [B*512, 768] activation: the [CLS] rows of the
    # extraction path's last layer), at row counts below and above one tile; same arithmetic per row as the dense call
    for (Bq, K) in [(1, 768), (7, 768), (130, 768), (256, 3072), (300, 768)]:
        N, S = 768, 512
        a_all = _mk(Bq * S, K, "cuda", 0.5)
        r_all = _mk(Bq * S, N, "cuda")
        w = _mk(N, K, "cuda", 0.05)
        bias = torch.randn(N, device="cuda") * 0.5
        gamma = 1.0 + 0.2 * torch.randn(N, device="cuda")
        beta = 0.3 * torch.randn(N, device="cuda")
        a0, r0 = a_all.view(Bq, S, K)[:, 0], r_all.view(Bq, S, N)[:, 0]
        y = ops.linear_resid_ln(a0, w, bias, r0, gamma, beta)
        y_all = ops.linear_resid_ln(a_all, w, bias, r_all, gamma, beta)
        torch.cuda.synchronize()
        zref = _gemm_ref(a0.contiguous(), w, 0, 0) + bias + r0.float()
        yref = torch.nn.functional.layer_norm(zref, (N,), gamma, beta, 1e-12)
        res.append(_err_report(y, yref, f"ln_pitched_rows_{Bq}x{K}", 6e-2))
        dense0 = y_all.view(Bq, S, N)[:, 0]
        res.append({"case": f"ln_pitched_rows_same_as_dense_{Bq}x{K}", "ok": bool(torch.equal(y, dense0)),
                    "max_abs": float((y.float() - dense0.float()).abs().max())})
    # a row with a large common offset: the chunked (mean, M2) merge must not cancel catastrophically
    M, K, N = 256, 768, 768
    a = _mk(M, K, "cuda", 0.5)
    w = _mk(N, K, "cuda", 0.05)
    r = (torch.randn(M, N, device="cuda") + 40.0).bfloat16()
    gamma = torch.ones(N, device="cuda")
    beta = torch.zeros(N, device="cuda")
    zref = _gemm_ref(a, w, 0, 0) + r.float()
    yref = torch.nn.functional.layer_norm(zref, (N,), gamma, beta, 1e-12)
    y = ops.linear_resid_ln(a, w, None, r, gamma, beta)
    torch.cuda.synchronize()
    res.append(_err_report(y, yref, "ln_offset40", 0.3))   # z itself is rounded to bf16 at |z| ~ 40 (ulp 0.25)
    return res



def case_cls_head():
    """Sequence-classification head fwd / bwd vs torch autograd (fp32)."""
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(11)
    res = []
    for (B, L) in [(3, 5), (64, 2), (130, 32), (7, 1)]:
        pooled = torch.tanh(torch.randn(B, 768, device="cuda"))
        w = torch.randn(L, 768, device="cuda") * 0.05
        b = torch.randn(L, device="cuda") * 0.1
        labels = torch.randint(0, L, (B,), device="cuda")
        logits, rl = ops.cls_head(pooled, w, b, labels)
        pre = torch.atanh(pooled.double().clamp(-0.999999, 0.999999)).float().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        lr = torch.tanh(pre) @ wr.T + br
        loss = torch.nn.functional.cross_entropy(lr, labels)
        loss.backward()
        res.append(_err_report(logits, (pooled @ w.T + b), f"cls_logits_{B}x{L}", 1e-4))
        res.append(_err_report(rl.mean().reshape(1), torch.nn.functional.cross_entropy(pooled @ w.T + b, labels).reshape(1), f"cls_loss_{B}x{L}", 1e-4))
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        scale = torch.full((1,), 1.0 / B, device="cuda")
        dpre = ops.cls_pool_bwd(pooled, logits, labels, scale, w, dw, db)
        torch.cuda.synchronize()
        res.append(_err_report(dw, wr.grad, f"cls_dw_{B}x{L}", 1e-4))
        res.append(_err_report(db, br.grad, f"cls_db_{B}x{L}", 1e-4))
        res.append(_err_report(dpre, pre.grad, f"cls_dpre_{B}x{L}", 5e-3 * float(pre.grad.abs().max()) + 1e-6))   # dpre is stored in bf16
    return res



def case_dropout():
    """Training-mode dropout kernels vs torch fp32 with the SAME masks (numpy twin of csrc/stk_rng.cuh)."""
    import numpy as np
    import torch
    from oracle import dropout_oracle as do
    from stonkgs_b200 import ops
    torch.manual_seed(5)
    res = []
    # --- row kernels
    M = 1000
    d = ops.Drop(seed=0xC0FFEE, site=77, p=0.1)
    x = _mk(M, 768, "cuda")
    r = _mk(M, 768, "cuda")
    keep = torch.from_numpy(do.keep_mask(d.seed, d.site, M, 768, d.thr)).cuda()
    scale = 128.0 / (128 - d.thr)
    y = ops.dropout(x, d)
    res.append({"case": "drop_fwd_exact", "ok": bool(torch.equal(y, torch.where(keep, x.float() * scale, torch.zeros(1, device="cuda")).bfloat16())),
                "drop_rate": float(1 - keep.float().mean())})
    res.append({"case": "drop_rate_13_128", "ok": abs(float(1 - keep.float().mean()) - 13 / 128) < 2e-3})
    gamma = 1.0 + 0.2 * torch.randn(768, device="cuda")
    beta = 0.3 * torch.randn(768, device="cuda")
    yl, z, mean, rstd = ops.dropout_resid_ln(x, r, gamma, beta, d, save_for_backward=True)
    zref = torch.where(keep, x.float() * scale, torch.zeros(1, device="cuda")) + r.float()
    res.append(_err_report(z, zref, "drop_resid_z", 4e-2))
    res.append(_err_report(yl, torch.nn.functional.layer_norm(zref, (768,), gamma, beta, 1e-12), "drop_resid_ln_y", 6e-2))
    res.append(_err_report(mean, z.float().mean(1), "drop_resid_mean", 2e-3))
    y_inf = ops.dropout_resid_ln(x, r, gamma, beta, d)
    res.append(_err_report(y_inf, torch.nn.functional.layer_norm(zref, (768,), gamma, beta, 1e-12), "drop_resid_ln_y_nosave", 6e-2))
    # --- fused LayerNorm backward + dropout mask + bias column sums
    dy = _mk(M, 768, "cuda")
    zz = _mk(M, 768, "cuda")
    mu = zz.float().mean(1)
    rs = torch.rsqrt(zz.float().var(1, unbiased=False) + 1e-12)
    dg0, db0 = torch.zeros(768, device="cuda"), torch.zeros(768, device="cuda")
    dx_plain = ops.layernorm_bwd(dy, zz, gamma, mu, rs, dg0, db0)
    dg1, db1, dbias = torch.zeros(768, device="cuda"), torch.zeros(768, device="cuda"), torch.zeros(768, device="cuda")
    dx_f, dxm_f = ops.layernorm_bwd(dy, zz, gamma, mu, rs, dg1, db1, dbias=dbias, drop=d)
    res.append({"case": "ln_bwd_fused_dx_same", "ok": bool(torch.equal(dx_f, dx_plain))})
    res.append(_err_report(dg1, dg0, "ln_bwd_fused_dgamma", 1e-3 * float(dg0.abs().max())))
    res.append({"case": "ln_bwd_fused_mask_exact", "ok": bool(torch.equal(dxm_f != 0, keep & (dxm_f != 0)) and
                                                                 bool((dxm_f[~keep] == 0).all()))})
    # (dx_plain is the bf16-rounded dx; the fused kernel masks, scales and sums the fp32 values: one bf16 ulp / row apart)
    ref_m = torch.where(keep, dx_plain.float() * scale, torch.zeros(1, device="cuda"))
    res.append(_err_report(dxm_f, ref_m, "ln_bwd_fused_dxm", 8e-3 * float(ref_m.abs().max())))
    res.append(_err_report(dbias, ref_m.sum(0), "ln_bwd_fused_dbias", 2e-2 * float(ref_m.sum(0).abs().max())))
    dbias2 = torch.zeros(768, device="cuda")
    dx_n = ops.layernorm_bwd(dy, zz, gamma, mu, rs, dg1, db1, dbias=dbias2)
    res.append(_err_report(dbias2, dx_plain.float().sum(0), "ln_bwd_fused_dbias_nodrop", 2e-2 * float(dx_plain.float().sum(0).abs().max())))
    res.append({"case": "ln_bwd_fused_nodrop_dx_same", "ok": bool(torch.equal(dx_n, dx_plain))})
    # --- attention forward / backward with injected masks
    for (B, S, use_bias) in [(2, 256, False), (3, 512, True)]:
        da = ops.Drop(seed=1234 + S, site=5, p=0.1)
        qkv = (torch.randn(B * S, 2304, device="cuda") * 0.8).bfloat16()
        bias = None
        if use_bias:
            mask = torch.ones(B, S, dtype=torch.int64, device="cuda")
            mask[0, 100:256] = 0
            mask[2, 40:256] = 0
            bias = ops.mask_to_bias(mask)
        keep = torch.from_numpy(do.keep_mask(da.seed, da.site, B * 12 * S, S, da.thr)).cuda().view(B, 12, S, S)
        sc = 128.0 / (128 - da.thr)
        xq = qkv.float().clone().requires_grad_(True)
        q, k, v = xq.view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
        s_ = q @ k.transpose(-1, -2) * 0.125
        if bias is not None:
            s_ = s_ + bias[:, None, None, :]
        pr = torch.softmax(s_, -1)
        prd = torch.where(keep, pr * sc, torch.zeros(1, device="cuda"))
        oref = (prd @ v).permute(0, 2, 1, 3).reshape(B * S, 768)
        out, lse = ops.attention(qkv, bias, B, S, save_lse=True, drop=da)
        res.append(_err_report(out, oref, f"attn_drop_fwd_S{S}", 3e-2))
        res.append(_err_report(lse, torch.logsumexp(s_, -1), f"attn_drop_lse_S{S}", 1e-3))
        dout = (torch.randn(B * S, 768, device="cuda") * 0.5).bfloat16()
        oref.backward(dout.float())
        dqkv = ops.attention_bwd(qkv, bias, B, S, out, dout, lse, drop=da)
        torch.cuda.synchronize()
        for nm, sl in (("dq", slice(0, 768)), ("dk", slice(768, 1536)), ("dv", slice(1536, 2304))):
            res.append(_err_report(dqkv[:, sl], xq.grad[:, sl], f"attn_drop_bwd_{nm}_S{S}", 4e-2))
        out0 = ops.attention(qkv, bias, B, S)
        res.append({"case": f"attn_drop_differs_S{S}", "ok": bool((out.float() - out0.float()).abs().max() > 1e-2)})
    return res



def case_gemm_majors():
    """dgrad (B MN-major) and wgrad (A and B MN-major, split-K reduce-add)."""
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(3)
    res = []
    # dgrad: dx[M, Kin] = dy[M, Nout] @ W[Nout, Kin]  -> A = dy (K-major, K=Nout), B = W stored [K=Nout][N=Kin]
    Mt, Nout, Kin = 520, 768, 3072
    dy = _mk(Mt, Nout, "cuda", 0.5)
    W = _mk(Nout, Kin, "cuda", 0.05)
    ref = dy.float() @ W.float()
    res.append(_err_report(ops.gemm(dy, W, M=Mt, N=Kin, K=Nout, b_major=1, epilogue=ops.EPI_BIAS), ref, "dgrad_bf16", 3e-2))
    res.append(_err_report(ops.gemm(dy, W, M=Mt, N=Kin, K=Nout, b_major=1, epilogue=ops.EPI_F32), ref, "dgrad_f32", 2e-2))
    r = _mk(Mt, Kin, "cuda")
    res.append(_err_report(ops.gemm(dy, W, M=Mt, N=Kin, K=Nout, b_major=1, epilogue=ops.EPI_BIAS_RESID, resid=r),
                           ref + r.float(), "dgrad_resid", 4e-2))
    c = torch.randn(Mt, Kin, device="cuda")
    want = c + ref
    ops.gemm(dy, W, M=Mt, N=Kin, K=Nout, b_major=1, epilogue=ops.EPI_F32_ADD, out=c)
    res.append(_err_report(c, want, "dgrad_f32_add", 2e-2))
    # wgrad: dW[Nout, Kin] = dy^T[Nout, Mt] @ x[Mt, Kin] -> A = dy stored [K=Mt][M=Nout], B = x stored [K=Mt][N=Kin]
    for (Mt, Nout, Kin, split) in [(512, 768, 768, 1), (2048, 768, 3072, 4), (1000, 2304, 768, 3), (304, 768, 768, 2)]:
        dy = _mk(Mt, Nout, "cuda", 0.5)
        x = _mk(Mt, Kin, "cuda", 0.5)
        ref = dy.float().T @ x.float()
        res.append(_err_report(ops.gemm(dy, x, M=Nout, N=Kin, K=Mt, a_major=1, b_major=1, epilogue=ops.EPI_F32), ref,
                               f"wgrad_f32_{Mt}", 1e-2 * Mt ** 0.5 / 4))
        c = torch.zeros(Nout, Kin, device="cuda")
        ops.gemm(dy, x, M=Nout, N=Kin, K=Mt, a_major=1, b_major=1, epilogue=ops.EPI_F32_ADD, out=c, split_k=split)
        res.append(_err_report(c, ref, f"wgrad_splitk{split}_{Mt}", 1e-2 * Mt ** 0.5 / 4))
    return res


def case_gemm_ce():
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(4)
    res = []
    M, V, K = 304, 28996, 768
    h = _mk(M, K, "cuda", 1.0)
    W = _mk(V, K, "cuda", 0.05)
    labels = torch.randint(0, V, (M,), device="cuda").int()
    labels[0] = V - 1
    labels[1] = 0
    pitch = 2 * ((V + 255) // 256)
    part = torch.zeros(M, pitch, 2, device="cuda")
    tgt = torch.zeros(M, device="cuda")
    ops.gemm(h, W, M=M, N=V, K=K, epilogue=ops.EPI_CE_STATS, labels=labels, ce_partial=part, tgt_logit=tgt)
    lse, row_loss = ops.ce_finalize(part, tgt, M)
    logits = h.float() @ W.float().T
    res.append(_err_report(lse, torch.logsumexp(logits, -1), "ce_lse", 2e-3))
    res.append(_err_report(tgt, logits.gather(1, labels.long()[:, None])[:, 0], "ce_tgt_logit", 2e-3))
    res.append(_err_report(row_loss, torch.nn.functional.cross_entropy(logits, labels.long(), reduction="none"),
                           "ce_row_loss", 3e-3))
    # chunked variant with n_offset (two calls over vocabulary blocks)
    part2 = torch.zeros(M, pitch, 2, device="cuda")
    tgt2 = torch.zeros(M, device="cuda")
    cut = 256 * 50
    ops.gemm(h, W[:cut], M=M, N=cut, K=K, epilogue=ops.EPI_CE_STATS, labels=labels, ce_partial=part2, tgt_logit=tgt2)
    ops.gemm(h, W[cut:], M=M, N=V - cut, K=K, epilogue=ops.EPI_CE_STATS, labels=labels, ce_partial=part2,
             tgt_logit=tgt2, n_offset=cut)
    lse2, _ = ops.ce_finalize(part2, tgt2, M)
    res.append(_err_report(lse2, torch.logsumexp(logits, -1), "ce_lse_chunked", 2e-3))
    res.append(_err_report(tgt2, tgt, "ce_tgt_chunked", 1e-6))
    # dlogit
    scale = torch.tensor([1.0 / M], device="cuda")
    dl = torch.empty(M, (V + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")[:, :V]
    ops.gemm(h, W, M=M, N=V, K=K, epilogue=ops.EPI_CE_DLOGIT, labels=labels, lse=lse, scale_dev=scale, out=dl)
    ref = torch.softmax(logits, -1)
    ref[torch.arange(M), labels.long()] -= 1
    ref /= M
    res.append(_err_report(dl, ref, "ce_dlogit", 3e-5))
    return res


def _attn_ref(qkv, bias, B, S):
    import torch
    q, k, v = qkv.float().view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * 0.125
    if bias is not None:
        s = s + bias[:, None, None, :]
    p = torch.softmax(s, -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * S, 768)
    return o, torch.logsumexp(s, -1)


def case_attn():
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(5)
    res = []
    for (B, S, masked) in [(2, 256, False), (2, 512, True), (1, 128, True), (3, 384, True)]:
        qkv = _mk(B * S, 2304, "cuda", 1.0)
        bias = None
        if masked:
            m = torch.ones(B, S, dtype=torch.long, device="cuda")
            for b in range(B):
                m[b, 40 + 17 * b: S // 2] = 0
            bias = ops.mask_to_bias(m)
        out, lse = ops.attention(qkv, bias, B, S, save_lse=True)
        ref, lref = _attn_ref(qkv, bias, B, S)
        res.append(_err_report(out, ref, f"attn_fwd_S{S}_mask{int(masked)}", 2e-2))
        res.append(_err_report(lse, lref, f"attn_lse_S{S}", 2e-3))
    # general additive key bias (not only the 0 / finfo.min padding mask): scattered finite values, a masked island in
    # the middle of a key block, masked groups that start and end off the 32-key group boundaries
    B, S = 3, 512
    qkv = _mk(B * S, 2304, "cuda", 1.0)
    bias = torch.zeros(B, S, device="cuda")
    bias[0] = torch.randn(S, device="cuda") * 2.0
    bias[1, 100:131] = -1.7
    bias[1, 300:420] = torch.finfo(torch.float32).min
    bias[2, 5:37] = torch.finfo(torch.float32).min
    bias[2, 64:96] = torch.finfo(torch.float32).min
    bias[2, 511] = torch.finfo(torch.float32).min
    out, lse = ops.attention(qkv, bias, B, S, save_lse=True)
    ref, lref = _attn_ref(qkv, bias, B, S)
    res.append(_err_report(out, ref, "attn_fwd_general_bias", 2e-2))
    res.append(_err_report(lse, lref, "attn_lse_general_bias", 2e-3))
    # persistent paths: several items per CTA (forward: 296 CTAs, backward: 148), ragged padding masks, so that the
    # item-boundary hand-overs (deferred output, prefetched Q / K / V / bias, flat barrier parities) are exercised
    for (B, S) in [(17, 512), (41, 256), (30, 128), (11, 384)]:
        qkv = _mk(B * S, 2304, "cuda", 1.0)
        m = torch.ones(B, S, dtype=torch.long, device="cuda")
        lens = torch.randint(3, S // 2 + 1, (B,), device="cuda")
        m[:, : S // 2] = (torch.arange(S // 2, device="cuda")[None, :] < lens[:, None]).long()
        m[0] = 1                                          # one element without any masked key
        bias = ops.mask_to_bias(m)
        out, lse = ops.attention(qkv, bias, B, S, save_lse=True)
        ref, lref = _attn_ref(qkv, bias, B, S)
        res.append(_err_report(out, ref, f"attn_fwd_multi_item_B{B}_S{S}", 2e-2))
        res.append(_err_report(lse, lref, f"attn_lse_multi_item_B{B}_S{S}", 2e-3))
        dout = _mk(B * S, 768, "cuda", 1.0)
        dqkv = ops.attention_bwd(qkv, bias, B, S, out, dout, lse)
        gref = _attn_bwd_ref(qkv, bias, B, S, dout)
        res.append(_err_report(dqkv, gref, f"attn_bwd_multi_item_B{B}_S{S}", 0.03 * gref.abs().max().item()))
        # determinism across launches (no race between items)
        out2, lse2 = ops.attention(qkv, bias, B, S, save_lse=True)
        res.append({"case": f"attn_fwd_deterministic_B{B}_S{S}", "ok": bool(torch.equal(out, out2) and torch.equal(lse, lse2))})
    # first query tile(s) only (stk_attn_fwd_qrows): the computed rows are bit-identical to the full kernel's, the other
    # rows of the output are not written; several items per CTA at the larger batch
    for (B, S, q_rows) in [(3, 512, 128), (700, 512, 128), (5, 384, 256), (2, 256, 128)]:
        qkv = _mk(B * S, 2304, "cuda", 1.0)
        m = torch.ones(B, S, dtype=torch.long, device="cuda")
        lens = torch.randint(3, S // 2 + 1, (B,), device="cuda")
        m[:, : S // 2] = (torch.arange(S // 2, device="cuda")[None, :] < lens[:, None]).long()
        bias = ops.mask_to_bias(m)
        full = ops.attention(qkv, bias, B, S)
        part = torch.full((B * S, 768), 7.0, dtype=torch.bfloat16, device="cuda")
        ops.attention(qkv, bias, B, S, out=part, q_rows=q_rows)
        torch.cuda.synchronize()
        f3, p3 = full.view(B, S, 768), part.view(B, S, 768)
        res.append({"case": f"attn_qrows_same_B{B}_S{S}_q{q_rows}", "ok": bool(torch.equal(f3[:, :q_rows], p3[:, :q_rows]))})
        res.append({"case": f"attn_qrows_untouched_B{B}_S{S}_q{q_rows}", "ok": bool((p3[:, q_rows:] == 7.0).all())})
    # large, growing scores: later key blocks dominate -> exercises the lazy O rescale in TMEM
    B, S = 2, 512
    qkv = _mk(B * S, 2304, "cuda", 1.0)
    ramp = torch.linspace(0.5, 6.0, S, device="cuda").repeat(B)[:, None]
    qkv[:, 768:1536] = (qkv[:, 768:1536].float() * ramp).bfloat16()      # keys grow with position
    qkv[:, :768] = (qkv[:, :768].float() * 2.0).bfloat16()
    out, lse = ops.attention(qkv, None, B, S, save_lse=True)
    ref, lref = _attn_ref(qkv, None, B, S)
    res.append(_err_report(out, ref, "attn_fwd_rescale_path", 3e-2))
    res.append(_err_report(lse, lref, "attn_lse_rescale_path", 2e-2))
    dout = _mk(B * S, 768, "cuda", 1.0)
    dqkv = ops.attention_bwd(qkv, None, B, S, out, dout, lse)
    gref = _attn_bwd_ref(qkv, None, B, S, dout)
    res.append(_err_report(dqkv, gref, "attn_bwd_large_scores", 0.04 * gref.abs().max().item()))
    return res


def _time(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case_perf():
    import torch
    from stonkgs_b200 import ops
    res = []
    M = 65536
    for (N, K, epi, name) in [(2304, 768, ops.EPI_BIAS, "qkv"), (768, 768, ops.EPI_BIAS_RESID, "wo"),
                              (3072, 768, ops.EPI_BIAS_GELU, "ffn1"), (768, 3072, ops.EPI_BIAS_RESID, "ffn2")]:
        a = _mk(M, K, "cuda", 0.5)
        w = _mk(N, K, "cuda", 0.05)
        bias = torch.zeros(N, device="cuda")
        r = _mk(M, N, "cuda") if epi == ops.EPI_BIAS_RESID else None
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        ms = _time(lambda: ops.gemm(a, w, M=M, N=N, K=K, epilogue=epi, bias=bias, resid=r, out=out))
        ms_t = _time(lambda: torch.nn.functional.linear(a, w))
        res.append({"case": f"perf_gemm_{name}", "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "torch_ms": ms_t,
                    "torch_tflops": 2.0 * M * N * K / ms_t / 1e9, "ok": True})
    for (B, S) in [(128, 256), (128, 512)]:
        qkv = _mk(B * S, 2304, "cuda", 1.0)
        out = torch.empty(B * S, 768, dtype=torch.bfloat16, device="cuda")
        ms = _time(lambda: ops.attention(qkv, None, B, S, out=out))
        fl = 4.0 * B * 12 * S * S * 64
        q, k, v = qkv.view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
        ms_t = _time(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        res.append({"case": f"perf_attn_S{S}", "ms": ms, "tflops": fl / ms / 1e9, "torch_sdpa_ms": ms_t,
                    "torch_tflops": fl / ms_t / 1e9, "ok": True})
    x = _mk(131072, 768, "cuda")
    g = torch.ones(768, device="cuda")
    b = torch.zeros(768, device="cuda")
    y = torch.empty_like(x)
    ms = _time(lambda: ops.layernorm(x, g, b, out=y))
    res.append({"case": "perf_layernorm_131072", "ms": ms, "gbs": 2 * x.numel() * 2 / ms / 1e6, "ok": True})
    return res


CASES = {
    "elementwise": case_elementwise,
    "gemm_basic": case_gemm_basic,
    "gemm_epilogues": case_gemm_epilogues,
    "gemm_majors": case_gemm_majors,
    "cls_head": case_cls_head,
    "dropout": case_dropout,
    "gemm_ln": case_gemm_ln,
    "gemm_ce": case_gemm_ce,
    "attn": case_attn,
    "perf": case_perf,
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        name = sys.argv[2]
        try:
            out = CASES[name]()
        except Exception as e:  # noqa: BLE001
            import traceback
            out = [{"case": name, "ok": False, "exception": repr(e), "trace": traceback.format_exc()[-1500:]}]
        print("PROBE_JSON " + json.dumps(out))
        return
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True,
                               text=True, timeout=300)
            got = None
            for line in r.stdout.splitlines():
                if line.startswith("PROBE_JSON "):
                    got = json.loads(line[len("PROBE_JSON "):])
            if got is None:
                got = [{"case": name, "ok": False, "rc": r.returncode, "stdout": r.stdout[-1500:], "stderr": r.stderr[-2500:]}]
        except subprocess.TimeoutExpired as e:
            got = [{"case": name, "ok": False, "timeout": True, "stdout": (e.stdout or b"")[-1000:].decode(errors="replace")
                    if isinstance(e.stdout, bytes) else str(e.stdout)[-1000:]}]
        for g in got:
            g["group"] = name
        results += got
        print(f"[{name}] {time.time() - t0:.1f}s", flush=True)
        for g in got:
            brief = {k: v for k, v in g.items() if k in ("case", "ok", "max_abs", "mean_abs", "ms", "tflops",
                                                         "torch_tflops", "gbs", "exception", "timeout", "rc", "bad_frac")}
            print("   ", brief, flush=True)
            if not g.get("ok", False):
                for k in ("stderr", "trace", "stdout"):
                    if k in g:
                        print("     ", k, ":", str(g[k])[-1200:], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
            json.dump(results, f, indent=1)
    n_bad = sum(1 for g in results if not g.get("ok", False))
    print(f"probe: {len(results) - n_bad} ok, {n_bad} failed")



# ------------------------------------------------------------------------------------------------
# end-to-end cases (appended after the first bring-up run)
# ------------------------------------------------------------------------------------------------
def _attn_bwd_ref(qkv, bias, B, S, dout):
    import torch
    x = qkv.float().clone().requires_grad_(True)
    q, k, v = x.view(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * 0.125
    if bias is not None:
        s = s + bias[:, None, None, :]
    o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, 768)
    o.backward(dout.float())
    return x.grad


def case_attn_bwd():
    import torch
    from stonkgs_b200 import ops
    torch.manual_seed(6)
    res = []
    for (B, S, masked) in [(2, 512, True), (1, 128, False), (2, 256, True)]:
        qkv = _mk(B * S, 2304, "cuda", 1.0)
        bias = None
        if masked:
            m = torch.ones(B, S, dtype=torch.long, device="cuda")
            for b in range(B):
                m[b, 40 + 17 * b: S // 2] = 0
            bias = ops.mask_to_bias(m)
        out, lse = ops.attention(qkv, bias, B, S, save_lse=True)
        dout = _mk(B * S, 768, "cuda", 1.0)
        dqkv = ops.attention_bwd(qkv, bias, B, S, out, dout, lse)
        ref = _attn_bwd_ref(qkv, bias, B, S, dout)
        tol = 0.03 * ref.abs().max().item()
        res.append(_err_report(dqkv[:, :768], ref[:, :768], f"attn_bwd_dq_S{S}", tol))
        res.append(_err_report(dqkv[:, 768:1536], ref[:, 768:1536], f"attn_bwd_dk_S{S}", tol))
        res.append(_err_report(dqkv[:, 1536:], ref[:, 1536:], f"attn_bwd_dv_S{S}", tol))
    return res


def _load_case(name):
    import numpy as np
    import torch
    from oracle import weights
    from transformers import BertConfig
    from stonkgs_b200.model import STonKGsForPreTraining
    fix = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    L, B, n_kg, seed_w, seed_b, _ = [int(v) for v in fix["meta"]]
    sd = weights.make_state_dict(n_kg, L, seed_w)
    rows = weights.make_kg_table(n_kg, seed_w)
    cfg = BertConfig(vocab_size=28996, num_hidden_layers=L)
    model = STonKGsForPreTraining(None, cfg, rows)
    model.load_state_dict(sd, strict=True)
    model.eval().to("cuda")
    batch = {k: torch.from_numpy(fix[k]) for k in ("input_ids", "attention_mask", "token_type_ids", "masked_lm_labels",
                                                   "ent_masked_lm_labels", "next_sentence_labels")}
    return fix, sd, rows, model, batch


def case_e2e_fwd():
    import torch
    res = []
    for name in ("L2_B2_N997", "L12_B2_N997", "L2_B3_N3001_fullmask"):
        fix, sd, rows, model, batch = _load_case(name)
        with torch.no_grad():
            out = model(**batch, return_dict=True)
        torch.cuda.synchronize()
        r = [int(v) for v in fix["rows"]]
        res.append(_err_report(out.pooler_output.cpu(), torch.from_numpy(fix["pooler_output"]), f"{name}_pooler", 5e-2))
        res.append(_err_report(out.hidden_states[:, r].cpu().reshape(-1, 768),
                               torch.from_numpy(fix["sequence_output_rows"]).reshape(-1, 768), f"{name}_seq_rows", 8e-2))
        res.append(_err_report(out.loss.cpu().reshape(1), torch.from_numpy(fix["loss"]).reshape(1), f"{name}_loss",
                               2e-3 * float(fix["loss"])))
        parts = [float(p) for p in model._last_loss_parts]
        res.append({"case": f"{name}_loss_parts", "ok": True, "got": parts,
                    "ref": [float(fix["mlm_loss"]), float(fix["elm_loss"])]})
        res.append(_err_report(out.seq_relationship_logits.cpu(), torch.from_numpy(fix["seq_relationship_logits"]),
                               f"{name}_nsp_logits", 2e-2))
        # KG table special rows + bit-exact node rows
        ids = [int(v) for v in fix["kg_probe_ids"]]
        got = model.kg_table[torch.tensor(ids, device="cuda")].cpu()
        ref = torch.from_numpy(fix["kg_probe_rows"])
        special = [i for i, v in enumerate(ids) if v in (100, 102, 103)]
        normal = [i for i, v in enumerate(ids) if v not in (100, 102, 103)]
        res.append({"case": f"{name}_kg_rows_bitexact", "ok": bool(torch.equal(got[normal], ref[normal]))})
        res.append(_err_report(got[special], ref[special], f"{name}_kg_special_rows", 5e-2))
    return res


def case_e2e_bwd():
    import torch
    from oracle import stonkgs_oracle as orc
    res = []
    for name in ("L2_B2_N997", "L12_B2_N997"):
        fix, sd, rows, model, batch = _load_case(name)
        model.zero_grad(set_to_none=True)
        out = model(**batch)
        loss = out[0]
        loss.backward()
        torch.cuda.synchronize()
        table = orc.build_kg_table(sd, rows)
        o, grads = orc.forward_backward(sd, table, batch)
        res.append(_err_report(loss.detach().cpu().reshape(1), o["loss"].detach().reshape(1), f"{name}_train_loss",
                               2e-3 * float(o["loss"])))
        named = dict(model.named_parameters())
        worst = []
        for k, g in grads.items():
            p = named[k]
            if p.grad is None:
                res.append({"case": f"{name}_grad_missing_{k}", "ok": False})
                continue
            got = p.grad.detach().cpu().float()
            denom = g.abs().max().item() + 1e-12
            rel = (got - g).abs().max().item() / denom
            cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1), g.reshape(1, -1)).item()
            worst.append((rel, cos, k))
        worst.sort(reverse=True)
        bad = [(round(r, 4), round(c, 5), k) for r, c, k in worst if (r > 0.08 and "key.bias" not in k)]
        res.append({"case": f"{name}_grads", "ok": len(bad) == 0, "n": len(worst), "bad": bad[:12],
                    "worst5": [(round(r, 4), round(c, 5), k) for r, c, k in worst[:5]],
                    "min_cos": min(c for r, c, k in worst if "key.bias" not in k)})
        dead = [k for k, p in named.items() if p.requires_grad and p.grad is None]
        res.append({"case": f"{name}_dead_params", "ok": sorted(dead) == sorted(str(x) for x in fix["dead_names"]),
                    "dead": dead})
    return res


def case_elm_head_perf():
    """BASELINE configs[3]: ELM head alone, ~1M entity nodes, batch 128 -> 4 864 labelled rows."""
    import torch
    from stonkgs_b200 import training
    N, R, H = 1_000_003, 128 * 38, 768
    W = (torch.randn(N, H, device="cuda") * 0.05).bfloat16()
    t = torch.randn(R, H, device="cuda").bfloat16()
    labels = torch.randint(0, N, (R,), device="cuda", dtype=torch.int32)
    ms_f = _time(lambda: training._ce_forward(t, W, labels), iters=5, warm=2)
    lse, _ = training._ce_forward(t, W, labels)
    dT = torch.zeros(R, H, dtype=torch.float32, device="cuda")
    gW = torch.zeros(N, H, dtype=torch.float32, device="cuda")
    scale = torch.full((1,), 1.0 / R, device="cuda")
    ms_b = _time(lambda: training._ce_backward(t, W, labels, lse, scale, dT, gW), iters=3, warm=1)
    fl = 2.0 * R * H * N
    return [{"case": "elm_head_1M_fwd", "ms": ms_f, "tflops": fl / ms_f / 1e9, "ok": True},
            {"case": "elm_head_1M_bwd", "ms": ms_b, "tflops": 3 * fl / ms_b / 1e9, "ok": True},
            {"case": "elm_head_1M_step", "ms": ms_f + ms_b, "tflops": 4 * fl / (ms_f + ms_b) / 1e9,
             "rows_per_s": R / (ms_f + ms_b) * 1e3, "ok": True}]


CASES.update({"attn_bwd": case_attn_bwd, "e2e_fwd": case_e2e_fwd, "e2e_bwd": case_e2e_bwd,
              "elm_head_perf": case_elm_head_perf})

if __name__ == "__main__":
    main()

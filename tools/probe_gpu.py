"""First-contact GPU probe: exercises the tcgen05 GEMM in every operand-major combination and the full engine steps
against plain torch fp32 on the same GPU.  Each case runs in its own subprocess so that a trapped kernel cannot poison
the rest.  Usage: python tools/probe_gpu.py [case ...]   (no args: run all cases in subprocesses)"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def split_planes(t, planes):
    import torch
    hi = t.to(torch.bfloat16)
    if planes == 1:
        return hi.contiguous(), hi.float()
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous(), hi.float() + lo.float()


def case_gemm(a_major, b_major, M, N, K, planes, splits):
    import torch
    from physicsvae_b200.engine import gemm_bf16
    torch.manual_seed(0)
    dev = torch.device("cuda")
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    Ap, Ar = split_planes(A if a_major == 0 else A.t().contiguous(), planes)
    Bp, Br = split_planes(B if b_major == 0 else B.t().contiguous(), planes)
    if a_major: Ar = Ar.t()
    if b_major: Br = Br.t()
    ref = (Ar.double() @ Br.double().t()).float()
    D = gemm_bf16(Ap, Bp, M, N, K, a_major, b_major, planes, splits)
    torch.cuda.synchronize()
    err = (D - ref).abs()
    tol = (1e-2 if planes == 1 else 2e-3) + 1e-3 * ref.abs()   # bf16x3 drops lo*lo: ~K * 2^-18
    bad = err > tol
    out = {"max_err": float(err.max()), "ref_absmax": float(ref.abs().max()), "bad_frac": float(bad.float().mean())}
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        out["bad_rows"] = [int(rows.min()), int(rows.max()), int(rows.numel())]
        out["bad_cols"] = [int(cols.min()), int(cols.max()), int(cols.numel())]
        out["D00"] = D[:2, :4].tolist(); out["R00"] = ref[:2, :4].tolist()
        out["nan"] = bool(torch.isnan(D).any())
    out["ok"] = not bool(bad.any())
    return out


def make_model(dsb, da, z, te, md, wm, dev):
    import torch
    torch.manual_seed(1)
    def mlp(i, hidden, o):
        dims = [i] + hidden + [o]
        Ws, bs = [], []
        for l in range(len(dims) - 1):
            w = torch.randn(dims[l + 1], dims[l], device=dev)
            w = w * ((1.0 if l < len(dims) - 2 else 0.3) / w.pow(2).sum(1, keepdim=True).sqrt())
            Ws.append(w.contiguous()); bs.append((0.1 * torch.randn(dims[l + 1], device=dev)).contiguous())
        return Ws, bs
    return {"task_encoder": mlp(2 * dsb, te, 2 * z), "motor_decoder": mlp(dsb + z, md, da),
            "world_model": mlp(dsb + da, wm, dsb), "value_branch": mlp(2 * dsb, te, 1)}


def ref_mlp(x, Ws, bs):
    import torch
    for l, (w, b) in enumerate(zip(Ws, bs)):
        x = torch.nn.functional.linear(x, w, b)
        if l < len(Ws) - 1:
            x = torch.relu(x)
    return x


def case_engine(dsb, da, z, te, md, wm, B, N, cursor, precision):
    import torch
    from physicsvae_b200.engine import Engine
    dev = torch.device("cuda")
    params = make_model(dsb, da, z, te, md, wm, dev)
    spec = {k: [(w.shape[0], "relu" if l < len(v[0]) - 1 else "linear") for l, w in enumerate(v[0])] for k, v in params.items()}
    eng = Engine(dsb, da, z, spec, latent_prior=True, precision=precision, max_batch=max(B, 128))
    grads = {}
    for k, (Ws, bs) in params.items():
        g = torch.zeros(eng.grad_elems(k), device=dev) if k != "value_branch" else None
        grads[k] = g
        eng.bind_net(k, Ws, bs, g)
    eng.sync_weights()
    torch.manual_seed(2)
    s = torch.randn(N + 1, dsb, device=dev, dtype=torch.float64)
    X = torch.cat([s[:-1], s[:-1] + 0.05 * s[1:]], 1)
    Y = (torch.rand(N, da, device=dev) * 2 - 1)
    eng.alloc_transitions(N)
    eng.ingest(X, Y)
    eng.set_cursor(cursor)
    x = X[cursor:cursor + B].float(); y = Y[cursor:cursor + B]
    if precision == "bf16":
        x = x.bfloat16().float(); y = y.bfloat16().float()
    s1, s2 = x[:, :dsb], x[:, dsb:]
    res = {}
    def flat_grads(Ws, bs):
        return torch.cat([torch.cat([w.grad.flatten(), b.grad.flatten()]) for w, b in zip(Ws, bs)])
    def cmp(name, got, ref, scale=1.0):
        err = float((got - ref).abs().max()) * scale
        rel = float((got - ref).norm() / (ref.norm() + 1e-30))
        res[name] = {"max_abs_err_scaled": err, "rel_l2": rel, "ref_norm": float(ref.norm())}
        return rel
    # ---- world step
    for k in params:
        for t in params[k][0] + params[k][1]:
            t.requires_grad_(True); t.grad = None
    Ws, bs = params["world_model"]
    pred = ref_mlp(torch.cat([s1, y], 1), Ws, bs)
    loss = torch.nn.functional.mse_loss(pred, s2)
    loss.backward()
    l = eng.world_step(B).clone()
    torch.cuda.synchronize()
    res["world_loss"] = [float(l[0]), float(loss)]
    worst = cmp("world_grad", grads["world_model"], flat_grads(Ws, bs), B)
    # ---- vae step
    for k in params:
        for t in params[k][0] + params[k][1]:
            t.grad = None
    eps = torch.randn(B, z, device=dev)
    h = ref_mlp(x, *params["task_encoder"])
    mu, lv = h[:, :z], h[:, z:]
    zt = mu + eps * torch.exp(0.5 * lv)
    a = ref_mlp(torch.cat([s1, zt], 1), *params["motor_decoder"])
    fut = ref_mlp(torch.cat([s1, a], 1), Ws, bs)
    la = torch.nn.functional.mse_loss(a, y)
    lk = torch.mean(-0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp(), dim=1), dim=0)
    lc = torch.nn.functional.mse_loss(fut, s2)
    kl_c, cyc_c = 1.0, 0.05
    tot = la + kl_c * lk + cyc_c * lc
    tot.backward()
    l = eng.vae_step(B, eps=eps, kl_coeff=kl_c, cyc_coeff=cyc_c).clone()
    torch.cuda.synchronize()
    res["vae_loss"] = [l[:5].tolist(), [float(tot), float(la), float(lk), 0.0, float(lc)]]
    worst = max(worst, cmp("te_grad", grads["task_encoder"], flat_grads(*params["task_encoder"]), B))
    worst = max(worst, cmp("md_grad", grads["motor_decoder"], flat_grads(*params["motor_decoder"]), B))
    # ---- forward API
    with torch.no_grad():
        o = eng.forward(x, 15, eps=eps, noise=True)
        torch.cuda.synchronize()
        worst = max(worst, cmp("fwd_action", o["action"], a))
        worst = max(worst, cmp("fwd_future", o["future"], fut))
        worst = max(worst, cmp("fwd_mu", o["mu"], mu))
        worst = max(worst, cmp("fwd_value", o["value"], ref_mlp(x, *params["value_branch"])[:, 0]))
    tol = 2e-2 if precision == "bf16" else 2e-4
    res["worst_rel_l2"] = worst
    res["ok"] = bool(worst < tol and abs(res["world_loss"][0] - res["world_loss"][1]) < tol * abs(res["world_loss"][1]))
    return res


CASES = {}
for am in (0, 1):
    for bm in (0, 1):
        CASES["gemm_min_%d%d" % (am, bm)] = lambda am=am, bm=bm: case_gemm(am, bm, 128, 256, 64, 1, 1)
        CASES["gemm_rag_%d%d" % (am, bm)] = lambda am=am, bm=bm: case_gemm(am, bm, 328, 200, 248, 1, 1)
        CASES["gemm_big_%d%d" % (am, bm)] = lambda am=am, bm=bm: case_gemm(am, bm, 4096, 1024, 1024, 1, 1)
        CASES["gemm_x3_%d%d" % (am, bm)] = lambda am=am, bm=bm: case_gemm(am, bm, 520, 456, 392, 2, 1)
CASES["gemm_split_11"] = lambda: case_gemm(1, 1, 1024, 1024, 8192, 1, 8)
CASES["gemm_split_x3_11"] = lambda: case_gemm(1, 1, 248, 1024, 5000 // 8 * 8, 2, 8)
CASES["gemm_n208_01"] = lambda: case_gemm(0, 1, 256, 208, 256, 1, 1)
CASES["engine_small_x3"] = lambda: case_engine(37, 11, 8, [48, 48], [64, 64, 64], [96, 96], 200, 1000, 300, "bf16x3")
CASES["engine_default_x3"] = lambda: case_engine(197, 45, 32, [256, 256], [512, 512, 512], [1024, 1024], 300, 1024, 256, "bf16x3")
CASES["engine_default_bf16"] = lambda: case_engine(197, 45, 32, [256, 256], [512, 512, 512], [1024, 1024], 1000, 2048, 1024, "bf16")

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        for name in sys.argv[1:]:
            t0 = time.time()
            try:
                r = CASES[name]()
            except Exception as e:  # noqa
                r = {"ok": False, "exception": repr(e)[:500]}
            r["case"] = name; r["sec"] = round(time.time() - t0, 2)
            print("PROBE " + json.dumps(r), flush=True)
        sys.exit(0)
    names = list(CASES)
    groups = [[n for n in names if n.startswith("gemm_min")], [n for n in names if n.startswith("gemm_rag")],
              [n for n in names if n.startswith("gemm_big")], [n for n in names if n.startswith("gemm_x3")],
              ["gemm_split_11"], ["gemm_split_x3_11"], ["engine_small_x3"], ["engine_default_x3"], ["engine_default_bf16"]]
    env = dict(os.environ)
    for g in groups:
        # one subprocess per case inside the group if the group run fails
        try:
            p = subprocess.run([sys.executable, __file__] + g, capture_output=True, text=True, timeout=240, env=env)
            out = p.stdout + p.stderr[-3000:]
        except subprocess.TimeoutExpired as e:
            out = "TIMEOUT %s\n%s" % (g, (e.stdout or b"")[-2000:] if isinstance(e.stdout, bytes) else e.stdout)
        print(out, flush=True)
        done = [l for l in out.splitlines() if l.startswith("PROBE ")]
        if len(done) < len(g):
            for n in g[len(done):][1:]:
                try:
                    p = subprocess.run([sys.executable, __file__, n], capture_output=True, text=True, timeout=120, env=env)
                    print(p.stdout + p.stderr[-2000:], flush=True)
                except subprocess.TimeoutExpired:
                    print("TIMEOUT", n, flush=True)
    env["PVAE_MN_BN_ALIGN"] = "16"
    try:
        p = subprocess.run([sys.executable, __file__, "gemm_n208_01", "gemm_rag_01", "gemm_rag_11"], capture_output=True, text=True, timeout=120, env=env)
        print("MN_BN_ALIGN=16:\n" + p.stdout + p.stderr[-2000:], flush=True)
    except subprocess.TimeoutExpired:
        print("TIMEOUT bn16", flush=True)

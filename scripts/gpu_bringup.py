"""GPU bring-up: run every piece of the hot path against the CPU oracle and print the errors.
Development aid (the asserted versions live in tests/test_gpu_*.py)."""
import json, math, os, sys, time, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import solver as S, cnn as OC
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
import safetensors.torch as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")

def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()

def section(name):
    print(f"\n=== {name} ===", flush=True)

def run(name, fn):
    try:
        fn()
    except Exception:
        print(f"[FAIL] {name}")
        traceback.print_exc()
    sys.stdout.flush()

def solver_case(nx, ny, batch):
    cfg = S.SolverConfig(nx=nx, ny=ny)
    u64, v64 = S.filtered_velocity_field(cfg, batch, seed=42)
    plan = Plan(PlanConfig(nx=nx, ny=ny, batch=batch))
    u, v = u64.float().to(dev), v64.float().to(dev)
    # explicit terms
    du, dv = plan.explicit_terms(u, v)
    odu, odv = S.explicit_terms(u64, v64, cfg, cfg.dt)
    o32u, o32v = S.explicit_terms(u64.float(), v64.float(), cfg, cfg.dt)
    print(f"[{nx}x{ny} b{batch}] F: K32 vs O64 {rel(du, odu):.3e} {rel(dv, odv):.3e} | O32 vs O64 {rel(o32u, odu):.3e} {rel(o32v, odv):.3e}")
    # projection on a non-solenoidal field
    g = torch.Generator().manual_seed(1)
    ru = torch.randn(batch, nx, ny, generator=g, dtype=torch.float64); rv = torch.randn(batch, nx, ny, generator=g, dtype=torch.float64)
    pu, pv = ru.float().to(dev), rv.float().to(dev)
    p = plan.project(pu, pv)
    ou, ov, op = S.pressure_projection(ru, rv, cfg)
    o32 = S.pressure_projection(ru.float(), rv.float(), cfg)
    d = plan.divergence(pu, pv)
    print(f"   P: u {rel(pu, ou):.3e} v {rel(pv, ov):.3e} p {rel(p, op):.3e} | O32 u {rel(o32[0], ou):.3e} p {rel(o32[2], op):.3e} | div_inf*h/vmax {d.abs().max().item()*cfg.hx/pu.abs().max().item():.3e}")
    # rk4 steps
    uu, vv = u.clone(), v.clone()
    o_u, o_v = u64.clone(), v64.clone()
    q_u, q_v = u64.float(), v64.float()
    inv64 = S.inverse_eigenvalues(cfg); inv32 = S.inverse_eigenvalues(cfg, torch.float32)
    for n in (1, 10):
        k = n - (0 if n == 1 else 1)
        plan.step(uu, vv, nsteps=k)
        for _ in range(k):
            o_u, o_v, o_p = S.rk4_step(o_u, o_v, cfg, None, inv64)
            q_u, q_v, _ = S.rk4_step(q_u, q_v, cfg, None, inv32)
        print(f"   step n={n}: K32 vs O64 {rel(uu, o_u):.3e} {rel(vv, o_v):.3e} | O32 vs O64 {rel(q_u, o_u):.3e} {rel(q_v, o_v):.3e}")
    pp = plan.new_field()
    plan.step(uu, vv, nsteps=1, p=pp)
    o_u, o_v, o_p = S.rk4_step(o_u, o_v, cfg, None, inv64)
    print(f"   step n=11 with p: u {rel(uu, o_u):.3e} p {rel(pp, o_p):.3e}")
    plan.close()

def cnn_case(nx, ny, batch, which="mlacfd", scale=1.0):
    cfgj = json.load(open(os.path.join(ROOT, "tests/golden", "MLACFD.json" if which == "mlacfd" else "cnn1.json")))
    sd = st.load_file(os.path.join(ROOT, "tests/golden", "mlacfd_weights.safetensors" if which == "mlacfd" else "cnn1_bn_weights.safetensors"))
    model = models.buildModel(cfgj); model.load_state_dict(sd)
    plan = Plan(PlanConfig(nx=nx, ny=ny, batch=batch))
    model.attach(plan)
    g = torch.Generator().manual_seed(7)
    u64 = torch.randn(batch, nx, ny, generator=g, dtype=torch.float64) * 3
    v64 = torch.randn(batch, nx, ny, generator=g, dtype=torch.float64) * 3
    u, v = u64.float().to(dev), v64.float().to(dev)
    y = plan.cnn_forward(u, v, scale)
    torch.cuda.synchronize()
    x = torch.stack((u64.float().double(), v64.float().double()), 1) * scale
    y_emu = OC.model_forward(x, cfgj, sd, round_bf16=True)
    y_64 = OC.model_forward(x, cfgj, sd, round_bf16=False)
    print(f"[cnn {which} {nx}x{ny} b{batch} s={scale:.3f}] K vs O_bf16emu {rel(y, y_emu):.3e} | O_bf16emu vs O64 {rel(y_emu, y_64):.3e} | max|y| {y.abs().max().item():.3e} finite {bool(torch.isfinite(y).all())}")
    # edges: compare first/last rows & columns separately
    yc, ye = y.double().cpu(), y_emu
    for nm, sl in (("x0", (slice(None), slice(None), 0)), ("xN", (slice(None), slice(None), -1)), ("y0", (slice(None), slice(None), slice(None), 0)), ("yN", (slice(None), slice(None), slice(None), -1))):
        print(f"      edge {nm}: {((yc[sl]-ye[sl]).norm()/ye[sl].norm()).item():.3e}", end="")
    print()
    uu, vv = u.clone(), v.clone()
    plan.cnn_correct(uu, vv, scale)
    print(f"   fused add: {rel(uu - u, y_emu[:, 0]):.3e} {rel(vv - v, y_emu[:, 1]):.3e}")
    plan.close()

def rollout_case(nx, ny, batch, nsteps):
    cfgj = json.load(open(os.path.join(ROOT, "tests/golden/MLACFD.json")))
    sd = st.load_file(os.path.join(ROOT, "tests/golden/mlacfd_weights.safetensors"))
    model = models.buildModel(cfgj); model.load_state_dict(sd)
    cfg = S.SolverConfig(nx=nx, ny=ny)
    u64, v64 = S.filtered_velocity_field(cfg, batch, seed=42)
    plan = Plan(PlanConfig(nx=nx, ny=ny, batch=batch)); model.attach(plan)
    u, v = u64.float().to(dev), v64.float().to(dev)
    t = time.time(); plan.step(u, v, nsteps=nsteps, apply_cnn=True); torch.cuda.synchronize(); t = time.time() - t
    ou, ov, _ = OC.rollout(u64, v64, cfg, nsteps, cfgj, sd, 1.0, round_bf16=True)
    fu, fv, _ = OC.rollout(u64, v64, cfg, nsteps, cfgj, sd, 1.0, round_bf16=False)
    print(f"[rollout {nx}x{ny} b{batch} n={nsteps}] K vs O(bf16emu cnn) {rel(u, ou):.3e} {rel(v, ov):.3e} | K vs O64 {rel(u, fu):.3e} | O(bf16emu) vs O64 {rel(ou, fu):.3e} | {t*1e3/nsteps:.3f} ms/step")
    hu, hv = u64.float().clone(), v64.float().clone()
    plan.rollout_host(hu, hv, nsteps, apply_cnn=True)
    print(f"   host rollout vs device rollout: {rel(hu, u):.3e}")
    plan.close()

def timing(nx, ny, batch):
    cfgj = json.load(open(os.path.join(ROOT, "tests/golden/MLACFD.json")))
    model = models.buildModel(cfgj)
    plan = Plan(PlanConfig(nx=nx, ny=ny, batch=batch)); model.attach(plan)
    names = ["explicit", "div_rfft", "col_solve", "irfft_grad", "conv_first", "conv_mid", "conv_last", "proj_cluster"]
    cells = nx * ny * batch
    for w, nm in enumerate(names):
        try:
            ms = plan.time_kernel(w, 3, 10)
        except Exception as e:
            print(f'   [{nx}x{ny} b{batch}] {nm:11s} n/a ({e})'); continue
        extra = ""
        if w == 5:
            extra = f" -> {2*9*64*64*cells/ms/1e9:.1f} TFLOP/s"
        print(f"   [{nx}x{ny} b{batch}] {nm:11s} {ms:8.4f} ms{extra}")
    u = torch.zeros(batch, nx, ny, device=dev); v = torch.zeros_like(u)
    for cnn in (False, True):
        plan.step(u, v, 2, apply_cnn=cnn); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); plan.step(u, v, 5, apply_cnn=cnn); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"   [{nx}x{ny} b{batch}] step cnn={cnn}: {ms:.3f} ms -> {cells/ms/1e6:.2f} Gcell-steps/s")
    plan.close()

if __name__ == "__main__":
    what = sys.argv[1:] or ["solver", "cnn", "rollout", "timing"]
    print(torch.cuda.get_device_name(0))
    if "solver" in what:
        section("solver")
        for c in ((64, 64, 1), (128, 128, 2), (64, 128, 3), (256, 256, 2), (512, 256, 1)):
            run(f"solver{c}", lambda c=c: solver_case(*c))
    if "cnn" in what:
        section("cnn")
        run("cnn64", lambda: cnn_case(64, 64, 1))
        run("cnn128", lambda: cnn_case(128, 128, 2))
        run("cnn256", lambda: cnn_case(256, 256, 3, scale=1 / 7))
        run("cnn1bn", lambda: cnn_case(64, 128, 2, which="cnn1"))
    if "rollout" in what:
        section("rollout")
        run("ro64", lambda: rollout_case(64, 64, 1, 10))
        run("ro128", lambda: rollout_case(128, 128, 2, 5))
    if "timing" in what:
        section("timing")
        run("t256", lambda: timing(256, 256, 64))
        run("t1024", lambda: timing(1024, 1024, 16))
    for w in what:
        if w.startswith("t="):
            nx_, ny_, b_ = (int(x) for x in w[2:].split(","))
            run(w, lambda: timing(nx_, ny_, b_))
    if "t128" in what:
        section("timing 128^2 x 512 (config 4 share of one of 8 GPUs)")
        run("t128", lambda: timing(128, 128, 512))

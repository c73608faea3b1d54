"""Measurement legs bench.py folds into its one JSON line (BASELINE.json configs C1, C2, C3, C5; the autograd-API timing;
the CPU baseline variants; the TF32 peak).  Everything here is timed with CUDA events on the current stream after warm-up,
or with perf_counter for the CPU legs.  Nothing here is the headline number."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

IMG_RES = 224
A_MANO = 31068.0              # algorithmic bytes / hand, MANO head fwd+bwd
F_GEMM = 2 * (46680 + 630180 + 298752)   # algorithmic FLOP / hand of the dense contractions, fwd + bwd (shape,
                                         # pose blend, LBS weight blend and their three transposes); MMA passes issued are not credited


def cuda_time(fn, reps, warm, dev):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) * 1e-3 / reps


def measure_tf32_peak(dev, n=8192, reps=10):
    """cuBLAS TF32 GEMM (torch.matmul, allow_tf32) n^3, best of `reps` -- the tensor-roofline denominator BASELINE.md asks the
    builder to measure (MEASURED_PEAKS.json holds bf16 only)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return 2.0 * n ** 3 / best / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def graphed(fn, dev):
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    return g.replay


def config_c2(dev, hbm_gbs, tf32_tflops, sizes=(1024, 65536)):
    """C2: MANOHead fwd+bwd (grads on v3d.cam, j3d.cam, j2d.norm), one hand side, through the C ABI with preallocated buffers."""
    from hands_b200.step import GeometryStep

    out = {}
    for B in sizes:
        st = GeometryStep(B, dev, with_pcl=False, hands_per_sample=1)

        def mano():
            st.mano_forward(0)
            st.mano_backward(0)

        t = cuda_time(mano, 30 if B <= 8192 else 8, 3, dev)
        r = {"ms": t * 1e3, "hands_per_s": B / t, "hbm_frac": B / t * A_MANO / (hbm_gbs * 1e9),
             "tensor_frac": (B / t * F_GEMM / (tf32_tflops * 1e12)) if tf32_tflops else None}
        if B <= 1024:
            tg = cuda_time(graphed(mano, dev), 50, 3, dev)
            r["cuda_graph_replay"] = {"ms": tg * 1e3, "hands_per_s": B / tg}
        out[f"B={B}"] = r
        del st
    return out


def config_c3(dev, hbm_gbs, sizes=(1024, 16384)):
    """C3: PerspectiveCropLayer fwd+bwd, n crops of 3x224x224, one crop per source image."""
    from hands_b200.step import GeometryStep

    out = {}
    for n in sizes:
        st = GeometryStep(n, dev, with_mano=False, hands_per_sample=1)
        st.pcl_setup()

        def pcl():
            st.pcl_forward()
            st.pcl_backward()

        t = cuda_time(pcl, 10, 2, dev)
        by = n * (3 * IMG_RES * IMG_RES * 4 * 3 + 12 * st.mean_s2)
        out[f"crops={n}"] = {"ms": t * 1e3, "crops_per_s": n / t, "alg_bytes": by, "hbm_frac": by / t / (hbm_gbs * 1e9)}
        del st
    return out


def autograd_api(dev, S=1024, reps=10):
    """The path the reference's training step would call: `perspective_crop` + `MANOHead` r/l as torch.autograd.Function
    drop-ins (outputs and workspaces allocated per call, autograd bookkeeping included), fwd + backward."""
    from hands_b200.pcl import perspective_crop
    from hands_b200.src.nets.hand_heads.mano_head import MANOHead
    from hands_b200.synthetic import synthetic_head_inputs, synthetic_pcl_inputs

    n = 2 * S
    _, bbox, Kc = synthetic_pcl_inputs(n, seed=5, img_res=IMG_RES, smin=IMG_RES // 4, smax=3 * IMG_RES // 4)
    bbox_h, Kc = bbox, Kc.to(dev)      # boxes arrive on the host, like the data loader's
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.randn(S, 3, IMG_RES, IMG_RES, generator=g, device=dev)
    g_crop = torch.randn(n, 3, IMG_RES, IMG_RES, generator=g, device=dev)
    heads, inp, gw = {}, {}, {}
    for side in (True, False):
        heads[side] = MANOHead(side, 1000.0, float(IMG_RES), synthetic=True).to(dev)
        inp[side] = [t.to(dev) for t in synthetic_head_inputs(S, seed=int(side))]
        gw[side] = (torch.randn(S, 778, 3, generator=g, device=dev), torch.randn(S, 21, 3, generator=g, device=dev), torch.randn(S, 21, 2, generator=g, device=dev))

    def step():
        x = img.requires_grad_(True)
        crop, rot = perspective_crop(x, bbox_h, Kc, img_res=IMG_RES, crops_per_img=2)
        rot = rot.view(S, 2, 3, 3)
        outs, grads, leaves = [crop], [g_crop], [x]
        for k, side in enumerate((True, False)):
            r, b, c, K = inp[side]
            r, b, c = r.requires_grad_(True), b.requires_grad_(True), c.requires_grad_(True)
            o = heads[side](r, b, c, K, pre_rot=rot[:, k].contiguous())
            pf = ".r" if side else ".l"
            outs += [o["v3d.cam" + pf], o["j3d.cam" + pf], o["j2d.norm" + pf]]
            grads += list(gw[side])
            leaves += [r, b, c]
        torch.autograd.backward(outs, grads)
        for t in leaves:
            t.grad = None

    t = cuda_time(step, reps, 3, dev)
    return {"samples": S, "ms": t * 1e3, "hands_per_s": 2 * S / t,
            "what": "perspective_crop + MANOHead.r/.l autograd.Function drop-ins, per-call allocation + autograd, fwd+bwd"}


def silhouette_consumer(dev, B=1024, reps=10):
    """§8(f3) consumer: MANORenderer (soft silhouette of mano.v3d.cam, faces_per_pixel = 10) + fused L1 mask loss, fwd + bwd,
    on the hand-sized synthetic mesh with MANO's counts (778 vertices, 1538 faces).  Issue-bound, not HBM-bound: the
    algorithmic bytes (vertices in, one fp32 mask out, its gradient in, vertex gradients out) are reported for scale only."""
    from hands_b200.losses import render_loss
    from hands_b200.src.models.hands_light.renderer import MANORenderer
    from hands_b200.synthetic import synthetic_silhouette_inputs

    vc, faces, K = synthetic_silhouette_inputs(B, seed=0, img_res=IMG_RES)
    r = MANORenderer({"img_res": IMG_RES}, faces_r=faces.numpy(), faces_l=faces.numpy()).to(dev)
    v, meta = vc.to(dev), {"intrinsics": K.to(dev)}
    gt = (torch.rand(B, 1, IMG_RES, IMG_RES, generator=torch.Generator().manual_seed(1)) > 0.5).float().to(dev)
    valid = torch.ones(B, device=dev)
    state = {}

    def fwd():
        state["x"] = v.requires_grad_(True)
        state["m"] = r({"mano.v3d.cam.r": state["x"]}, meta, is_right=True)["mask"]

    def full():
        fwd()
        render_loss(state["m"], gt, valid).backward()
        state["x"].grad = None

    t_f = cuda_time(fwd, reps, 3, dev)
    t = cuda_time(full, reps, 3, dev)
    cov = float((state["m"] > 0.5).float().mean())
    abytes = B * (2 * 778 * 12 + 2 * IMG_RES * IMG_RES * 4)
    return {"meshes": B, "fwd_ms": t_f * 1e3, "fwd_bwd_ms": t * 1e3, "meshes_per_s": B / t, "coverage": cov,
            "algorithmic_gbytes_per_s": abytes / t / 1e9,
            "what": "MANORenderer soft silhouette (778 verts, 1538 faces, 224^2, 10 faces/pixel) + L1 mask loss, fwd+bwd"}


# ---- CPU baselines ----------------------------------------------------------------------------------------------------
class CpuReferenceStep:
    """Reference torch CPU path for the C4 step: PCL (grid_sample + interpolate per crop, as the reference closure does) +
    orientation fix-up + MANOHead right/left, forward and backward.  `crop_range` restricts the PCL part to a slice of the
    crops (used by the worker-pool variant)."""

    def __init__(self, samples, seed=0, with_mano=True, crop_range=None):
        from hands_b200.synthetic import synthetic_head_inputs, synthetic_mano_buffers, synthetic_pcl_inputs
        from oracle import geometry_oracle as O

        self.O, self.S, self.with_mano = O, samples, with_mano
        n = samples * 2
        g = torch.Generator().manual_seed(seed)
        _, self.bbox, self.Kc = synthetic_pcl_inputs(n, seed=seed, img_res=IMG_RES, smin=IMG_RES // 4, smax=3 * IMG_RES // 4)
        self.img = torch.randn(samples, 3, IMG_RES, IMG_RES, generator=g)
        self.g_crops = torch.randn(n, 3, IMG_RES, IMG_RES, generator=g)
        self.crop_range = crop_range
        self.hands = []
        if with_mano:
            for side in range(2):
                rotmat, betas, cam, K = synthetic_head_inputs(samples, seed=seed + 10 * side)
                self.hands.append(dict(buf=synthetic_mano_buffers(side == 0), rotmat=rotmat, betas=betas, cam=cam, K=K,
                                       g_v3d=torch.randn(samples, 778, 3, generator=g), g_j3d=torch.randn(samples, 21, 3, generator=g),
                                       g_j2d=torch.randn(samples, 21, 2, generator=g)))

    def _pcl_per_sample(self, lo, hi):
        """Crops [lo, hi) forward + backward, one source image at a time as the reference's data loader does (each image its
        own autograd leaf: slicing one big leaf would make every crop's backward allocate a batch-sized zero gradient)."""
        O = self.O
        acc = 0.0
        rots = []
        b = lo
        while b < hi:
            im = b // 2
            e = min(hi, 2 * im + 2)
            x = self.img[im : im + 1].clone().requires_grad_(True)
            crops, rot = O.perspective_crop(x.expand(e - b, -1, -1, -1), self.bbox[b:e], self.Kc[b:e], IMG_RES)
            torch.autograd.backward([crops], [self.g_crops[b:e]])
            acc += float(x.grad[0, 0, 0, 0])
            rots.append(rot)
            b = e
        return acc, torch.cat(rots)

    def run_pcl_slice(self):
        lo, hi = self.crop_range
        return self._pcl_per_sample(lo, hi)[0]

    def _mano(self, rot):
        O = self.O
        outs, grads, leaves = [], [], []
        for side, h in enumerate(self.hands):
            r = h["rotmat"].clone().requires_grad_(True)
            b = h["betas"].clone().requires_grad_(True)
            c = h["cam"].clone().requires_grad_(True)
            pose = O.pcl_fix_global_orient(rot[:, side], r) if rot is not None else r
            o = O.mano_head_forward(h["buf"], pose, b, c, h["K"], float(IMG_RES), 0.1)
            outs += [o["v3d.cam"], o["j3d.cam"], o["j2d.norm"]]
            grads += [h["g_v3d"], h["g_j3d"], h["g_j2d"]]
            leaves += [r, b, c]
        torch.autograd.backward(outs, grads)
        return sum(float(x.grad.sum()) for x in leaves)

    def run(self, batched_s=None):
        if batched_s is None:
            acc, rot = self._pcl_per_sample(0, 2 * self.S)
        else:
            img = self.img.clone().requires_grad_(True)
            crops, rot = batched_pcl_fixed_s(self.O, img.repeat_interleave(2, dim=0), self.bbox, self.Kc, batched_s)
            torch.autograd.backward([crops], [self.g_crops])
            acc = float(img.grad.sum())
        return acc + self._mano(rot.view(self.S, 2, 3, 3))


def batched_pcl_fixed_s(O, src, bbox, K, s):
    """The batched form SURVEY.md 8(d) names: ONE grid_sample (n, s, s) + ONE interpolate for the whole batch.  torch can only
    batch crops of one intermediate size, so every box is given side `s` here (the mean side) -- an upper bound on what a
    batched CPU implementation of the reference could deliver, not the reference's per-sample loop."""
    import torch.nn.functional as F

    n = src.shape[0]
    Ps, rots = [], []
    bb = bbox.clone()
    cx, cy = (bb[:, 0] + bb[:, 2]) // 2, (bb[:, 1] + bb[:, 3]) // 2
    bb[:, 0], bb[:, 2], bb[:, 1], bb[:, 3] = cx - s // 2, cx - s // 2 + s, cy - s // 2, cy - s // 2 + s
    for q in range(n):
        P, Rv, _ = O.pcl_homography(bb[q].tolist(), K[q], IMG_RES)
        Ps.append(P)
        rots.append(Rv)
    P = torch.stack(Ps).float()
    lin = torch.linspace(0, 1, s)
    u, v = lin.view(1, 1, s).expand(n, s, s), lin.view(1, s, 1).expand(n, s, s)
    pts = torch.stack([u, v, torch.ones_like(u)], dim=-1)                     # (n, s, s, 3), row j <-> v, column i <-> u
    xyz = torch.einsum("nab,nijb->nija", P, pts)
    grid = xyz[..., :2] / (1e-8 + xyz[..., 2:3]) / IMG_RES * 2 - 1
    mid = F.grid_sample(src, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    out = F.interpolate(mid, size=(IMG_RES, IMG_RES), mode="bilinear", align_corners=True)
    return out, torch.stack(rots).float()


def _time_cpu(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / max(steps, 1)


def cpu_loop_all_threads(samples, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    st = CpuReferenceStep(samples)
    dt = _time_cpu(st.run, steps, warmup)
    return 2 * samples / dt, dt


def cpu_loop_one_thread(samples, steps=1, warmup=0):
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        st = CpuReferenceStep(samples)
        dt = _time_cpu(st.run, steps, warmup)
    finally:
        torch.set_num_threads(nt)
    return 2 * samples / dt, dt


def cpu_batched(samples, steps, warmup, s=112):
    torch.set_num_threads(os.cpu_count() or 1)
    st = CpuReferenceStep(samples)
    dt = _time_cpu(lambda: st.run(batched_s=s), steps, warmup)
    return 2 * samples / dt, dt


_POOL_STATE = {}


def _pool_init():
    import torch as _t

    _t.set_num_threads(1)


def _pool_prepare(args):
    seed, samples, lo, hi = args
    st = CpuReferenceStep(samples, seed=seed, with_mano=False, crop_range=(lo, hi))
    st.run_pcl_slice()
    _POOL_STATE["st"] = st
    return os.getpid()


def _pool_run(_):
    return _POOL_STATE["st"].run_pcl_slice()


def cpu_worker_pool(samples, steps, warmup, workers=None, timeout=240):
    """How the reference itself parallelises the crop layer: one single-threaded process per DataLoader worker
    (src/datasets/hands_light_dataset.py runs inside torch DataLoader workers), here `workers` processes each handling a
    contiguous slice of the step's crops fwd+bwd, while the parent runs the batched MANO heads on all threads.  Returns
    (hands/s, seconds per step, workers) or None if the pool could not be had in time."""
    import multiprocessing as mp

    workers = workers or min(os.cpu_count() or 1, 2 * samples, 32)   # (each spawned worker imports torch: bounded start-up time and memory)
    n = 2 * samples
    bounds = [(w * n // workers, (w + 1) * n // workers) for w in range(workers)]
    ctx = mp.get_context("spawn")
    try:
        pools = [ctx.Pool(1, initializer=_pool_init) for _ in range(workers)]   # one process per slice so each keeps its own state
    except Exception:
        return None
    try:
        prep = [p.apply_async(_pool_prepare, ((0, samples, lo, hi),)) for p, (lo, hi) in zip(pools, bounds)]
        for r in prep:
            r.get(timeout=timeout)
        torch.set_num_threads(os.cpu_count() or 1)
        parent = CpuReferenceStep(samples)

        def one():
            rs = [p.apply_async(_pool_run, (0,)) for p in pools]
            parent._mano(None)
            for r in rs:
                r.get(timeout=timeout)

        dt = _time_cpu(one, steps, warmup)
        return 2 * samples / dt, dt, workers
    except Exception:
        return None
    finally:
        for p in pools:
            p.terminate()


def config_c1():
    """C1: MANO layer forward, right hand, batch 64, torch.no_grad(), reference torch path on the CPU (the oracle)."""
    from hands_b200.synthetic import synthetic_mano_buffers
    from oracle import geometry_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    buf = synthetic_mano_buffers(True)
    g = torch.Generator().manual_seed(0)
    pose, betas = 0.3 * torch.randn(64, 48, generator=g), torch.randn(64, 10, generator=g)
    with torch.no_grad():
        dt = _time_cpu(lambda: O.mano_forward(buf, betas, pose[:, :3], pose[:, 3:]), 20, 3)
    return {"ms": dt * 1e3, "hands_per_s": 64 / dt, "cores": os.cpu_count() or 1, "what": "oracle.mano_forward (smplx restatement), B=64, no_grad, CPU torch"}

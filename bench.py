#!/usr/bin/env python
"""bench.py — RMCL kernels-only training step on B200 (BASELINE.json configs[1]).

One *step* = one pass of the hot path over one 256-sample batch on one GPU:
    momentum EMA of the 161-tensor / 111.7 M-parameter ViLT-B/32 key encoder (fp32 master params)
 -> fused InfoNCE forward+backward of q[256,256] against [k ; queue[256,65536] bf16], tau 0.07
 -> (N>1) exchange of the normalised keys + ring-buffer enqueue of the gathered keys: one fused kernel over NVLink peer
    memory (rmcl_gather_enqueue_p2p), or ncclAllGather + the enqueue kernel (--exchange nccl / no symmetric memory)
 -> (N=1) ring-buffer enqueue of the keys.
Data-parallel weak scaling: every rank runs that step on its own batch, the queue is replicated
and every rank enqueues the identical gathered keys; ``value`` = (rank-steps all ranks completed)
/ (max-over-ranks device time).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA)
    python bench.py --impl reference [...]                          # reference arithmetic on host cores

The JSON line carries ``roofline`` (dominant kernel, CUDA-event timed inside this run),
``kernels`` (every kernel of the step, same arithmetic; at N=1 also the path's other kernels at their BASELINE shapes:
cfg3 PGD updates, the cfg4-shaped and cfg5 InfoNCE calls, the Barlow-Twins loss), ``cpu_baseline`` (oracle port timed on
this box's host cores, rank 0 at N=1), ``e2e`` (host buffers -> C-ABI ``rmcl_step_host`` -> host
results, copies inside the timed region), ``clocks`` and ``gpu_launches``.
Only the cpu_baseline / --impl reference legs touch oracle/ (as the thing being timed there is
the reference arithmetic itself); the CUDA arm never imports it.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=256, C=256, K=65536, tau=0.07, m=0.999)
WORKLOAD = "cfg2: RMCL kernels-only step (EMA 161 tensors/111.7M fp32 params + fused InfoNCE fwd+bwd B256 C256 K65536 bf16 queue + enqueue)"
METRIC = "RMCL steps/s (PGD+MoCo InfoNCE) at 1/2/4/8 B200; kernel % of roofline"


def load_shapes():
    path = os.path.join(ROOT, "tests", "golden", "vilt_b32_key_encoder_shapes.txt")
    shapes = []
    for line in open(path):
        line = line.strip()
        if line and not line.startswith("#"):
            shapes.append(tuple(int(d) for d in line.split("x")))
    return shapes


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


def traffic_table():
    """dram bytes per launch from the committed ncu --set full captures (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(path)) if os.path.isfile(path) else {}


# ---------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # torch device index -> physical index when CUDA_VISIBLE_DEVICES remaps
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001 - no NVML: report nulls rather than fail the bench
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except AttributeError:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------ reference arm
def cpu_step_factory(sample_frac, seed=0):
    """Builds the CPU leg: the oracle's kernels-only step (the reference's own ATen expressions,
    fp32 — the reference has no bf16 CPU path) on a fraction of the workload."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rmcl_oracle as O
    torch.manual_seed(seed)
    shapes = load_shapes()
    if sample_frac < 1.0:
        total = sum(int(torch.Size(s).numel()) for s in shapes)
        keep, acc = [], 0
        for s in sorted(shapes, key=lambda s: torch.Size(s).numel()):   # small tensors first: keeps the per-tensor loop cost
            n = torch.Size(s).numel()
            if acc + n <= total * sample_frac or not keep:
                keep.append(s)
                acc += n
        shapes = keep
        ema_frac = acc / total
    else:
        ema_frac = 1.0
    B, C = CFG["B"], CFG["C"]
    K = max(B, int(CFG["K"] * sample_frac) // B * B)
    pk = [torch.randn(s) for s in shapes]
    pq = [torch.randn(s) for s in shapes]
    q = torch.randn(B, C)
    k_hat = O.l2_normalize(torch.randn(B, C))
    state = {"queue": torch.randn(C, K), "ptr": 0, "pk": pk}

    def step():
        new_k, res, new_queue, new_ptr = O.rmcl_kernel_step(state["pk"], pq, CFG["m"], q, k_hat, state["queue"],
                                                            state["ptr"], CFG["tau"])
        state.update(pk=new_k, queue=new_queue, ptr=new_ptr)
        return float(res["loss"])

    return step, dict(K=K, ema_frac=ema_frac, infonce_frac=K / CFG["K"])


def time_cpu(steps, warmup, budget_s):
    """Times the CPU leg; shrinks the per-step sample so that the run fits ``budget_s``.
    Returns (full-workload steps/s, description)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frac = 1.0
    step, info = cpu_step_factory(frac)
    t0 = time.perf_counter()
    step()
    t_first = time.perf_counter() - t0
    need = t_first * (steps + warmup)
    if need > budget_s:
        frac = max(1.0 / 64, budget_s / need)
        step, info = cpu_step_factory(frac)
        step()
    for _ in range(max(0, warmup - 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    # a sampled step does (ema_frac, infonce_frac) of the work; both legs are linear in their size
    scale = 0.5 * (info["ema_frac"] + info["infonce_frac"]) if frac < 1.0 else 1.0
    full_dt = dt / scale
    sample = (f"{steps} steps of the oracle port (reference ATen expressions, fp32, torch CPU {cores} threads) on "
              f"{'the full cfg2 step' if frac >= 1.0 else 'a %.3f sample of cfg2 (K=%d, %.3f of EMA params), scaled linearly' % (frac, info['K'], info['ema_frac'])}")
    return 1.0 / full_dt, dt * 1e3, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, ms, cores, sample = time_cpu(args.steps, max(1, args.warmup), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, **CFG},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA arm has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import rmcl_b200  # noqa: F401  (raises if librmcl_b200.so is missing)
    from rmcl_b200 import ops

    B, C, K, tau, m = CFG["B"], CFG["C"], CFG["K"], CFG["tau"], CFG["m"]
    pk_peak = peaks()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    shapes = load_shapes()
    n_params = sum(torch.Size(s).numel() for s in shapes)
    params_k = [torch.randn(s, device=dev, generator=g) for s in shapes]
    params_q = [torch.randn(s, device=dev, generator=g) for s in shapes]
    plan = ops.EmaPlan(params_k, params_q)
    gq = torch.Generator(device=dev).manual_seed(7)       # queue identical on every rank (replicated)
    queue = torch.randn(C, K, device=dev, generator=gq).bfloat16()
    ptr = torch.zeros(1, dtype=torch.int64, device=dev)
    q = torch.randn(B, C, device=dev, generator=g).bfloat16()
    k_raw = torch.randn(B, C, device=dev, generator=g).bfloat16()
    gathered = torch.empty(world * B, C, dtype=torch.float32, device=dev) if world > 1 else None
    path = args.path

    # N>1: the key exchange (NCCL all-gather, latency-bound) and the enqueue of the gathered keys ride a
    # side stream under the HBM-bound EMA, which shares no data with them; the two streams are joined at
    # the end of every step, so a step stays one unit.  (The EMA has no data dependency on the InfoNCE
    # either in this kernels-only step — the link in the full step is the key-encoder forward — but
    # running those two concurrently measured no gain: both want every SM.)
    main_stream = torch.cuda.current_stream()
    side_stream = torch.cuda.Stream() if world > 1 else None
    ev_fwd, ev_side = torch.cuda.Event(), torch.cuda.Event()

    # N>1 exchange: the fused peer-memory kernel (rmcl_gather_enqueue_p2p: push over NVLink, signal, wait, enqueue — one
    # launch) or, if symmetric memory cannot be set up on this box / --exchange nccl, ncclAllGather + the enqueue kernel
    p2p = None
    if world > 1 and args.exchange in ("auto", "p2p"):
        try:
            from rmcl_b200.dist import P2PKeyExchange
            p2p = P2PKeyExchange(B, C, dev)
        except Exception as e:      # noqa: BLE001
            if args.exchange == "p2p":
                raise
            sys.stderr.write(f"[bench] peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather\n")

    def exchange_and_enqueue(keys, impl=None):
        if world > 1 and p2p is not None and impl != "nccl":
            p2p.enqueue_(queue, keys, ptr)
            return
        if world > 1:
            dist.all_gather_into_tensor(gathered, keys)
            keys = gathered
        ops.enqueue_(queue, keys, ptr)

    def step(qq=None, kk=None, after_fwd=None):
        qq, kk = (q, k_raw) if qq is None else (qq, kk)
        if world == 1:
            ops.ema_multi_(plan, m)
            res = ops.infonce_fwd_bwd(qq, kk, queue, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
            exchange_and_enqueue(res["k_hat"])
            return res
        res = ops.infonce_fwd_bwd(qq, kk, queue, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        ev_fwd.record(main_stream)
        side_stream.wait_event(ev_fwd)
        with torch.cuda.stream(side_stream):
            if after_fwd is not None:
                after_fwd(res)          # e2e: the result read-back goes out as soon as the InfoNCE is done
            exchange_and_enqueue(res["k_hat"])
            ev_side.record(side_stream)
        ops.ema_multi_(plan, m)
        main_stream.wait_event(ev_side)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(3, args.warmup)):
        res = step()
    sampler = ClockSampler(local)
    sampler.start()
    total_ms = timed(step, args.steps)
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = world * args.steps / (total_ms * 1e-3)
    loss_val = float(res["loss"])

    # ---- end-to-end: host buffers -> C-ABI -> host results, every step
    q_host, k_host = q.cpu().pin_memory(), k_raw.cpu().pin_memory()
    if world == 1:
        host_step = ops.HostStep(plan, queue, ptr, B, C, tau, m, torch.bfloat16, path)
        h2d, d2h = host_step.h2d_bytes, host_step.d2h_bytes

        def e2e_step():
            host_step(q_host, k_host)
    else:
        q_dev, k_dev = torch.empty_like(q), torch.empty_like(k_raw)
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        dq_host = torch.empty(B, C, dtype=torch.float32).pin_memory()
        h2d, d2h = 2 * B * C * 2, 4 + B * C * 4

        def e2e_step():
            q_dev.copy_(q_host, non_blocking=True)
            k_dev.copy_(k_host, non_blocking=True)
            def read_back(r):
                loss_host.copy_(r["loss"], non_blocking=True)
                dq_host.copy_(r["dq"], non_blocking=True)
            step(q_dev, k_dev, after_fwd=read_back)
            torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps)
    e2e_value = world * args.steps / (e2e_ms * 1e-3)

    # ---- N>1: the exchange step alone, both implementations, back to back on every rank
    exchange_info = None
    if world > 1:
        keys_x = torch.nn.functional.normalize(torch.randn(B, C, device=dev, generator=g), dim=1)
        exchange_info = {"impl": "p2p" if p2p is not None else "nccl"}
        for impl in (("p2p", "nccl") if p2p is not None else ("nccl",)):
            for _ in range(5):
                exchange_and_enqueue(keys_x, impl)
            exchange_info[f"us_{impl}"] = timed(lambda: exchange_and_enqueue(keys_x, impl), 50) / 50 * 1000

    # ---- per-kernel durations inside the step (events on the launching stream, same order, so
    #      each kernel sees the cache state it sees in the real step: the EMA's 1.34 GB of traffic
    #      evicts the 33.5 MB queue from the 126 MB L2 before every InfoNCE pass)
    kern_ms = {"ema": 0.0, "infonce_prep": 0.0, "infonce_partial": 0.0, "infonce_finalize": 0.0, "enqueue": 0.0}
    reps = min(args.steps, 20)
    ops.profile_enable(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(reps):
        ev[0].record()
        ops.ema_multi_(plan, m)
        ev[1].record()
        r = ops.infonce_fwd_bwd(q, k_raw, queue, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        keys = r["k_hat"]
        if world > 1:
            dist.all_gather_into_tensor(gathered, keys)
            keys = gathered
        ev[2].record()
        ops.enqueue_(queue, keys, ptr)
        ev[3].record()
        torch.cuda.synchronize()
        st = ops.profile_infonce_ms()
        kern_ms["ema"] += ev[0].elapsed_time(ev[1]) / reps
        kern_ms["infonce_prep"] += st["prep"] / reps
        kern_ms["infonce_partial"] += st["partial"] / reps
        kern_ms["infonce_finalize"] += st["finalize"] / reps
        kern_ms["enqueue"] += ev[2].elapsed_time(ev[3]) / reps
    ops.profile_enable(False)

    flops_infonce = 4.0 * B * C * (K + 1)          # fused fwd (q.K^T) + bwd (P.K): 2 GEMMs of 2*B*C*K
    alg = {
        "ema": ("hbm", 12.0 * n_params),                                   # read k, read q, write k (fp32)
        "infonce_partial": ("tensor", flops_infonce),
        "enqueue": ("hbm", world * B * C * (4 + 2)),                      # read fp32 keys, write bf16 columns
    }
    traffic = traffic_table()
    kernels = {}
    for name, (bound, work) in alg.items():
        t = kern_ms[name] * 1e-3
        if bound == "hbm":
            ach, peak, unit = work / t / 1e9, pk_peak["hbm"], "GB/s"
        else:
            ach, peak, unit = work / t / 1e12, pk_peak["tf_sust"], "TFLOP/s"
        kernels[name] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                         "ms": kern_ms[name], "traffic": traffic.get(name)}
    # ---- the tcgen05 partial kernel over CONSECUTIVE launches (one event pair around the whole batch).
    #      An event pair around a single launch inside a PDL chain also times ~5-10 us of launch/event
    #      latency (tools/tc_timeline.py: %globaltimer span of the kernel 20-21 us, 3-4 us of it spent in
    #      griddepcontrol.wait for prep, against 27-30 us between the bracketing events), so the bracketed
    #      number above is an upper bound.  Here 8 distinct queue copies (8 x C*K*2 B > the 126 MB L2) are
    #      cycled so that every launch streams its queue from HBM as it does in the step.
    if world == 1 and path in ("auto", "tcgen05"):
        gq2 = torch.Generator(device=dev).manual_seed(11)
        n_copies = max(2, int(math.ceil(160e6 / (C * K * 2))) + 1)
        queues = [queue] + [torch.randn(C, K, device=dev, generator=gq2).bfloat16() for _ in range(n_copies - 1)]
        ops.infonce_fwd_bwd(q, k_raw, queue, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        n_b2b = 8 * n_copies

        def partial_batch():
            for j in range(n_b2b):
                ops.infonce_fwd_bwd(q, k_raw, queues[j % n_copies], tau, normalize_k=True, path=path,
                                    want=(), _partial_only=True)

        partial_batch()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()          # the launches are replayed as one graph so that the host (ctypes
        with torch.cuda.graph(graph):           # call overhead ~ kernel duration) cannot starve the stream
            partial_batch()
        graph.replay()
        ms_b2b = min(timed(graph.replay, 1) for _ in range(5)) / n_b2b
        kp = kernels["infonce_partial"]
        kp["ms_consecutive"] = ms_b2b
        kp["achieved_consecutive"] = flops_infonce / (ms_b2b * 1e-3) / 1e12
        kp["frac_consecutive"] = kp["achieved_consecutive"] / kp["peak"]
        kp["method"] = ("ms/achieved/frac: one CUDA-event pair around the single launch inside the step (includes launch + "
                        "event latency of a PDL-chained launch); *_consecutive: one event pair around %d back-to-back "
                        "launches (one CUDA graph) cycling over %d queue copies (> L2), divided by the count" % (n_b2b, n_copies))
        del graph, queues

    # ---- cfg3: the PGD update kernel alone (not part of the cfg2 step): B=128, 5 steps, pixel and
    #      embedding perturbations; 12 B/element (read g, read delta, write delta)
    if world == 1 and not args.no_pgd:
        for tag, shape, mode, lr, eps in (("pgd_pixel_ref_linf", (128, 3, 384, 384), "ref_linf", 0.05, 8 / 255),
                                          ("pgd_pixel_l2", (128, 3, 384, 384), "l2", 0.5, 1.0),
                                          ("pgd_embed_ref_linf", (128, 185, 768), "ref_linf", 0.05, 8 / 255),
                                          ("pgd_embed_sign_linf", (128, 185, 768), "sign_linf", 2 / 255, 8 / 255)):
            grad = torch.randn(shape, device=dev, generator=g)
            delta = torch.zeros(shape, device=dev)
            for _ in range(3):
                ops.pgd_step_(delta, grad, lr, eps, mode)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.pgd_step_(delta, grad, lr, eps, mode)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 5 * 1e-3
            nbytes = 12.0 * grad.numel()
            kernels[tag] = {"bound": "hbm", "achieved": nbytes / t / 1e9, "peak": pk_peak["hbm"], "unit": "GB/s",
                            "frac": nbytes / t / 1e9 / pk_peak["hbm"], "ms": t * 1e3, "traffic": traffic.get(tag),
                            "shape": list(shape)}
            del grad, delta
    # ---- cfg5 per-GPU InfoNCE (B512 C768 K262144 bf16: the two-pass tcgen05 variant, prep + S pass + PV pass +
    #      finalize timed as one call; queue 403 MB + P~ 268 MB > L2, so every call streams from HBM)
    if world == 1 and not args.no_pgd and path in ("auto", "tcgen05"):
        B5, C5, K5 = 512, 768, 262144
        g5 = torch.Generator(device=dev).manual_seed(5)
        q5 = torch.randn(B5, C5, device=dev, generator=g5).bfloat16()
        k5 = torch.randn(B5, C5, device=dev, generator=g5).bfloat16()
        queue5 = torch.nn.functional.normalize(torch.randn(C5, K5, device=dev, generator=g5), dim=0).bfloat16()
        for _ in range(3):
            ops.infonce_fwd_bwd(q5, k5, queue5, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        ms5 = timed(lambda: ops.infonce_fwd_bwd(q5, k5, queue5, tau, normalize_k=True, path=path,
                                                want=("loss", "dq", "k_hat")), 10) / 10
        f5 = 4.0 * B5 * C5 * (K5 + 1)
        kernels["infonce_cfg5_two_pass"] = {"bound": "tensor", "achieved": f5 / (ms5 * 1e-3) / 1e12, "peak": pk_peak["tf_sust"],
                                            "unit": "TFLOP/s", "frac": f5 / (ms5 * 1e-3) / 1e12 / pk_peak["tf_sust"], "ms": ms5,
                                            "traffic": traffic.get("infonce_cfg5_two_pass"), "shape": [B5, C5, K5],
                                            "launches": ["infonce_prep_kernel", "infonce_s_kernel", "infonce_pv_kernel",
                                                         "infonce_finalize_kernel"]}
        del q5, k5, queue5
    # ---- cfg4 per-GPU InfoNCE shape (B128 C128 K65536 bf16): arithmetic intensity 2*B = 256 flop per queue byte puts it
    #      at the HBM/tensor ridge; whole call (prep + tcgen05 partial + finalize) by CUDA-graph replay over queue copies
    #      larger than L2, so every call streams its 16.8 MB queue from HBM
    if world == 1 and not args.no_pgd and path in ("auto", "tcgen05"):
        B4, C4, K4 = 128, 128, 65536
        g4 = torch.Generator(device=dev).manual_seed(4)
        q4 = torch.randn(B4, C4, device=dev, generator=g4).bfloat16()
        k4 = torch.randn(B4, C4, device=dev, generator=g4).bfloat16()
        queues4 = [torch.nn.functional.normalize(torch.randn(C4, K4, device=dev, generator=g4), dim=0).bfloat16() for _ in range(12)]
        for qq in queues4[:2]:
            ops.infonce_fwd_bwd(q4, k4, qq, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for j in range(48):
                ops.infonce_fwd_bwd(q4, k4, queues4[j % 12], tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        graph.replay()
        ms4 = min(timed(graph.replay, 1) for _ in range(5)) / 48
        kernels["infonce_cfg4_b128_c128_call"] = {"bound": "hbm", "achieved": C4 * K4 * 2 / (ms4 * 1e-3) / 1e9, "peak": pk_peak["hbm"],
                                                  "unit": "GB/s", "frac": C4 * K4 * 2 / (ms4 * 1e-3) / 1e9 / pk_peak["hbm"], "ms": ms4,
                                                  "tflops": 4.0 * B4 * C4 * (K4 + 1) / (ms4 * 1e-3) / 1e12, "traffic": None,
                                                  "shape": [B4, C4, K4], "timing": "CUDA graph of 48 whole calls over 12 queue copies (> L2)"}
        del graph, queues4
    # ---- Barlow-Twins objective at the reference's size (batch 128, projector 8192) and at an 8-GPU gathered batch (1024):
    #      Gram formulation (default; 6*Bg^2*D flop, bound by reading q, k once and four short launches) and the direct
    #      D x D kernel (4*B*D^2 flop, tensor bound); each timed as one call = prep + tcgen05 kernel(s) + finalize
    if world == 1 and not args.no_pgd:
        gb = torch.Generator(device=dev).manual_seed(6)
        for tag, Bb, Db, bpath, bound in (("barlow_gram_b128_d8192", 128, 8192, "gram", "hbm"),
                                          ("barlow_direct_b128_d8192", 128, 8192, "direct", "tensor"),
                                          ("barlow_gram_b1024_d8192", 1024, 8192, "gram", "tensor")):
            kb_ = torch.randn(Bb, Db, device=dev, generator=gb)
            qb_ = 0.7 * kb_ + 0.7 * torch.randn(Bb, Db, device=dev, generator=gb)
            for _ in range(3):
                ops.barlow_fwd_bwd(qb_, kb_, 1.0 / Bb, 0.0051, path=bpath)
            graph = torch.cuda.CUDAGraph()      # graph replay: the host cost of 3-4 short launches would dominate a Python loop
            with torch.cuda.graph(graph):
                for _ in range(10):
                    ops.barlow_fwd_bwd(qb_, kb_, 1.0 / Bb, 0.0051, path=bpath)
            graph.replay()
            msb = min(timed(graph.replay, 1) for _ in range(5)) / 10
            if bound == "tensor":
                work = (6.0 * Bb * Bb * Db) if bpath == "gram" else (4.0 * Bb * Db * Db)
                ach, peak, unit = work / (msb * 1e-3) / 1e12, pk_peak["tf_sust"], "TFLOP/s"
            else:
                work = 3.0 * Bb * Db * 4         # read q, k (fp32), write dq
                ach, peak, unit = work / (msb * 1e-3) / 1e9, pk_peak["hbm"], "GB/s"
            kernels[tag] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "ms": msb,
                            "traffic": traffic.get(tag), "shape": [Bb, Db], "timing": "CUDA graph of 10 calls"}
            del graph, qb_, kb_
    for name in ("infonce_prep", "infonce_finalize"):
        kernels[name] = {"ms": kern_ms[name]}
    dominant = max(alg, key=lambda n: kern_ms[n])
    roofline = dict(kernels[dominant], kernel=dominant, peak_source=pk_peak["source"] +
                    (", sustained bf16" if kernels[dominant]["bound"] == "tensor" else ""))

    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, **CFG, "arithmetic": "InfoNCE: bf16 queue/q/k operands, fp32 accumulation and statistics; EMA: fp32 (bit-exact with ATen); enqueue: fp32 keys -> bf16 queue", "per_gpu_batch": B, "global_batch": world * B, "infonce_path": path,
                   "parallelism": f"dp{world}", "exchange": exchange_info,
                   "streams": "single stream" if world == 1 else "key exchange + enqueue on a side stream under the EMA, joined every step", "unit_of_value": "256-sample rank-steps per second, summed over ranks",
                   "l2": "inputs larger than L2: each step streams 1.34 GB of parameters (EMA) between InfoNCE passes; L2 is 126 MB"},
        "roofline": roofline, "kernels": kernels,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "api": "rmcl_step_host (C-ABI)" if world == 1 else "rmcl_b200.ops + NCCL all-gather"},
        "gpu_launches": 5 * args.steps, "launches_per_step": ["ema_multi_kernel", "infonce_prep_kernel",
                                                             "infonce_simt_kernel|infonce_tc_kernel", "infonce_finalize_kernel",
                                                             "gather_enqueue_p2p_kernel" if p2p is not None else "enqueue_kernel"],
        "clocks": clocks, "loss": loss_val,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample = time_cpu(steps=5, warmup=1, budget_s=25.0)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1 key exchange: fused peer-memory kernel (p2p) or ncclAllGather + enqueue (nccl); auto = p2p if available")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pgd", action="store_true", help="skip the cfg3 PGD-kernel and cfg5 InfoNCE roofline lines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()

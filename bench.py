#!/usr/bin/env python
"""bench.py — RMCL kernels-only training step on B200 (BASELINE.json configs[1]); ``--config cfg4`` = the full step.

One *step* (cfg2) = one pass of the hot path over one 256-sample batch on one GPU:
    momentum EMA of the 161-tensor / 111.7 M-parameter ViLT-B/32 key encoder (fp32 master params)
 -> fused InfoNCE forward+backward of q[256,256] against [k ; queue[256,65536] bf16], tau 0.07
 -> (N>1) exchange of the normalised keys + ring-buffer enqueue of the gathered keys: one fused kernel over NVLink peer
    memory (rmcl_gather_enqueue_p2p), or ncclAllGather + the enqueue kernel (--exchange nccl / no symmetric memory)
 -> (N=1) ring-buffer enqueue of the keys.
Data-parallel weak scaling: every rank runs that step on its own batch, the queue is replicated
and every rank enqueues the identical gathered keys; ``value`` = (rank-steps all ranks completed)
/ (max-over-ranks device time).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA)
    python bench.py --impl reference [...]                          # reference arithmetic on host cores, full cfg2 step
    python bench.py --config cfg4 [...]                             # BASELINE configs[3]: full RMCL step (tools/full_step.py)

Timing protocol: W warm-up steps, then EXACTLY K steps between a barrier + synchronize on both sides, CUDA events on the
launching stream, max over ranks.  A K-step block of this workload lasts only milliseconds, so the block is repeated
(each repeat bracketed the same way) until >= 0.25 s have been timed; ``ms_per_step`` is the MEDIAN block / K and the
min/max are reported beside it, the clock sampler runs over all of them.

The JSON line carries ``roofline`` (dominant kernel, CUDA-event timed inside this run), ``kernels`` (every kernel of the
step; at N=1 also the path's other kernels at their BASELINE shapes: cfg3 PGD updates, the cfg4-shaped and cfg5 InfoNCE
calls, the fp32-accurate InfoNCE, the Barlow-Twins loss), ``cpu_baseline`` (oracle port timed on this box's host cores,
rank 0 at N=1), ``gpu_baseline`` (the reference's expressions in eager torch and under torch.compile on the same GPU,
N=1), ``e2e`` (host buffers -> C-ABI ``rmcl_step_host`` -> host results, copies inside the timed region),
``parity_check`` (queue replicas identical on all ranks and equal to a replay through all_gather + enqueue), ``clocks``
and ``gpu_launches``.  Only the cpu_baseline / --impl reference legs touch oracle/ (the thing being timed there is the
reference arithmetic itself); the CUDA arm never imports it.
"""
import argparse
import hashlib
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=256, C=256, K=65536, tau=0.07, m=0.999)
WORKLOAD = "cfg2: RMCL kernels-only step (EMA 161 tensors/111.7M fp32 params + fused InfoNCE fwd+bwd B256 C256 K65536 bf16 queue + enqueue)"
METRIC = "RMCL steps/s (PGD+MoCo InfoNCE) at 1/2/4/8 B200; kernel % of roofline"
MIN_TIMED_S = 0.25


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


def source_hash():
    """sha256 over the CUDA sources and the header: ties profiles/traffic.json to the build it was captured from."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "robust-multimodal-contrastive-learning_b200", "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh")))
    files.append(os.path.join(ROOT, "include", "rmcl_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def traffic_table():
    """dram bytes per launch from this build's ncu --set full capture (profiles/traffic.json, written by
    tools/make_traffic.py together with the hash of the sources it was captured from)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(path):
        return {}, None
    t = json.load(open(path))
    meta = t.pop("_meta", {}) if isinstance(t.get("_meta"), dict) else {}
    cur = source_hash()
    return t, {"source_hash": meta.get("source_hash"), "current_source_hash": cur,
               "stale": meta.get("source_hash") != cur, "capture": meta.get("capture")}


def tensor_entry(flops, ms, pk):
    """Roofline entry of a tensor-bound kernel: kernels shorter than 1 ms are bursts, so ``frac`` is against the burst
    peak; the fraction of the sustained peak is kept beside it."""
    ach = flops / (ms * 1e-3) / 1e12
    burst = ms < 1.0
    peak = pk["tf_burst"] if burst else pk["tf_sust"]
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_kind": "burst" if burst else "sustained", "frac_of_sustained": ach / pk["tf_sust"],
            "frac_of_nominal_2250": ach / 2250.0, "ms": ms}


def hbm_entry(nbytes, ms, pk):
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "ms": ms}


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
def cpu_step_factory(seed=0):
    """The CPU leg: the oracle's kernels-only step (the reference's own ATen expressions, fp32 — the reference has no
    bf16 CPU path) on the FULL cfg2 workload: all 161 tensors, B256 C256 K65536."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rmcl_oracle as O
    torch.manual_seed(seed)
    shapes = load_shapes()
    B, C, K = CFG["B"], CFG["C"], CFG["K"]
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

    return step


def time_cpu(steps, warmup):
    """Times ``steps`` full cfg2 steps of the CPU leg after ``warmup`` (>= 2: the first touches ~2 GB of fresh pages)
    untimed ones.  No sampling, no scaling: the workload is the CUDA arm's.  Returns (steps/s, ms/step, cores, text)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_factory()
    for _ in range(max(2, warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = (f"{steps} steps of the oracle port (reference ATen expressions, fp32, torch CPU {cores} threads) on the full "
              f"cfg2 step (161 tensors / 111.7 M params, B256 C256 K65536), after {max(2, warmup)} warm-up steps")
    return 1.0 / dt, dt * 1e3, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == "cfg4":
        print(json.dumps({"impl": "reference", "unavailable": "cfg4 (full ViLT-B/32 step, batch 128) has no host-core arm: one "
                          "CPU step takes minutes; the reference arm is defined on the default workload (cfg2)"}), flush=True)
        return
    # a CPU step is ~0.1 s: cap the count so that the arm ends within a few minutes whatever K the driver passes
    steps = min(args.steps, 400)
    v, ms, cores, sample = time_cpu(steps, max(2, min(args.warmup, 10)))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": max(2, min(args.warmup, 10)), "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, **CFG},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU baseline (same GPU, torch)
def gpu_baseline(dev, shapes, q, k_raw, queue_bf16, steps, use_compile=True):
    """SURVEY 8(d) "GPU baseline to beat": the reference's own expressions for the cfg2 step, run by torch on this GPU —
    eagerly, as the reference executes them today, and under ``torch.compile``.  Restated from the reference lines
    (nothing from oracle/ is imported):
        EMA      objectives.py:219-224 (x4 at 257-260): per tensor  k.data = k.data*m + q.data*(1-m)
        InfoNCE  objectives.py:326-334+351 under autocast (Lightning precision=16 -> bf16 here): normalize, queue.clone(),
                 two einsums, cat, /T, CrossEntropyLoss on float logits, backward to q
        enqueue  objectives.py:244-248: int(ptr) (device sync), strided copy of keys.T, pointer write
    The queue is the fp32 buffer the reference holds (vilt_module.py:92); autocast casts it per call, as it does there."""
    import torch
    import torch.nn.functional as F
    B, C, K, tau, m = CFG["B"], CFG["C"], CFG["K"], CFG["tau"], CFG["m"]
    g = torch.Generator(device=dev).manual_seed(99)
    pk = [torch.randn(s, device=dev, generator=g) for s in shapes]
    pq = [torch.randn(s, device=dev, generator=g) for s in shapes]
    queue = queue_bf16.float()
    ptr = torch.zeros(1, dtype=torch.long, device=dev)
    q_raw = q.float()
    k_hat = F.normalize(k_raw.float(), dim=1)

    def ema():
        for a, b in zip(pk, pq):
            a.data = a.data * m + b.data * (1.0 - m)

    def infonce_loss(qr, kh, qu):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            qn = F.normalize(qr, dim=1)
            l_pos = torch.einsum("nc,nc->n", [qn, kh]).unsqueeze(-1)
            l_neg = torch.einsum("nc,ck->nk", [qn, qu.clone().detach()])
            logits = torch.cat([l_pos, l_neg], dim=1) / tau
        labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
        return F.cross_entropy(logits.float(), labels)

    def enqueue():
        p = int(ptr)
        queue[:, p:p + B] = k_hat.T
        ptr[0] = (p + B) % K

    def make_step(loss_fn, ema_fn):
        def step():
            ema_fn()
            qr = q_raw.detach().requires_grad_(True)
            loss = loss_fn(qr, k_hat, queue)
            loss.backward()
            enqueue()
            return loss
        return step

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {"unit": "steps/s", "workload": "the cfg2 step as the reference's torch expressions (fp32 queue buffer, bf16 autocast "
           "InfoNCE, per-tensor EMA loop, host-synchronising enqueue), same GPU, Python loop + CUDA events"}
    n = max(5, min(steps, 50))
    ms = timed(make_step(infonce_loss, ema), n)
    out["eager"] = {"value": 1e3 / ms, "ms_per_step": ms, "steps": n}
    if use_compile:
        try:
            t0 = time.perf_counter()
            c_loss = torch.compile(infonce_loss)

            def ema_out(ks, qs):
                return [a * m + b * (1.0 - m) for a, b in zip(ks, qs)]
            c_ema = torch.compile(ema_out)

            def ema_compiled():
                new = c_ema([a.data for a in pk], [b.data for b in pq])
                for a, v in zip(pk, new):
                    a.data = v
            ms = timed(make_step(c_loss, ema_compiled), n)
            out["torch_compile"] = {"value": 1e3 / ms, "ms_per_step": ms, "steps": n,
                                    "compile_s": round(time.perf_counter() - t0 - 3 * ms * 1e-3 - n * ms * 1e-3, 1),
                                    "what": "torch.compile (inductor, default mode) of the InfoNCE loss (forward+backward) and of "
                                            "the 161-tensor EMA expression; the enqueue keeps its int(ptr) sync"}
        except Exception as e:  # noqa: BLE001 - a baseline arm must not take the bench down
            out["torch_compile"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    return out


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

    # N>1: the key exchange (latency-bound) and the enqueue of the gathered keys ride a side stream under the
    # HBM-bound EMA, which shares no data with them.  The main stream joins the side stream after the EMA, i.e.
    # immediately before the next step's InfoNCE — the first consumer of the enqueued columns — so rank skew up to one
    # EMA duration is absorbed instead of being paid every step.  (The EMA has no data dependency on the InfoNCE
    # either in this kernels-only step — the link in the full step is the key-encoder forward — but running those two
    # concurrently measured no gain: both want every SM.)
    # The side stream has HIGH priority and the EMA is released only once the side stream has passed its wait for the
    # InfoNCE (ev_go): the EMA fills every SM with long-lived CTAs (2048 threads per SM, two waves of ~110 us), so an exchange
    # kernel that loses the race for the SMs would not start before the EMA's first wave retires — or, behind the EMA's own
    # second wave, not before its end, which serialises the exchange behind the EMA (measured: 0.333 ms/step at 8 GPUs).
    main_stream = torch.cuda.current_stream()
    side_stream = torch.cuda.Stream(priority=-1) if world > 1 else None
    ev_fwd, ev_side, ev_go = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

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
            ev_go.record(side_stream)
            if after_fwd is not None:
                after_fwd(res)          # e2e: the result read-back goes out as soon as the InfoNCE is done
            exchange_and_enqueue(res["k_hat"])
            ev_side.record(side_stream)
        main_stream.wait_event(ev_go)
        ops.ema_multi_(plan, m)
        main_stream.wait_event(ev_side)  # stream order: ... EMA(n) | join | InfoNCE(n+1): the join sits in front of its consumer
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

    def timed_blocks(fn, n):
        """K-step blocks, each bracketed as the contract says, repeated until MIN_TIMED_S has been timed (the repeat
        count is derived from the first block and agreed between the ranks)."""
        first = timed(fn, n)
        reps = max(1, min(500, int(math.ceil(MIN_TIMED_S * 1e3 / max(first, 1e-3)))))
        if world > 1:
            t = torch.tensor([reps], device=dev)
            dist.broadcast(t, 0)
            reps = int(t.item())
        blocks = [first] + [timed(fn, n) for _ in range(reps - 1)]
        return statistics.median(blocks), blocks

    warm = max(3, args.warmup)
    for _ in range(warm):
        res = step()
    sampler = ClockSampler(local)
    sampler.start()
    total_ms, blocks = timed_blocks(step, args.steps)
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = world * args.steps / (total_ms * 1e-3)
    loss_val = float(res["loss"])

    # ---- end-to-end: host buffers -> C-ABI -> host results, every step
    q_host, k_host = q.cpu().pin_memory(), k_raw.cpu().pin_memory()
    if world == 1:
        host_step = ops.HostStep(plan, queue, ptr, B, C, tau, m, torch.bfloat16, path)
        h2d, d2h = host_step.h2d_bytes, host_step.d2h_bytes
        e2e_api = "rmcl_step_host (C-ABI, one call per step: H2D q,k -> EMA -> InfoNCE -> enqueue -> D2H loss,dq)"

        def e2e_step():
            host_step(q_host, k_host)
    else:
        q_dev, k_dev = torch.empty_like(q), torch.empty_like(k_raw)
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        dq_host = torch.empty(B, C, dtype=torch.float32).pin_memory()
        h2d, d2h = 2 * B * C * 2, 4 + B * C * 4
        e2e_api = ("rmcl_b200.ops (C-ABI per op) with pinned-host q,k in and loss,dq out; key exchange = " +
                   ("rmcl_gather_enqueue_p2p (fused peer-memory kernel)" if p2p is not None else "ncclAllGather + rmcl_enqueue"))

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
    e2e_ms, e2e_blocks = timed_blocks(e2e_step, args.steps)
    e2e_value = world * args.steps / (e2e_ms * 1e-3)

    # ---- parity check of the state this run produced (driver-visible evidence for the N>1 path): from a common
    #      snapshot, run the real step() over a full ring wrap with keys that change every step; then restore the
    #      snapshot and replay the same keys through all_gather_into_tensor + the plain enqueue kernel (at N=1: the
    #      reference's strided-copy expression in eager torch).  Queue replicas must be bit-identical across ranks and
    #      equal to the replay; same for the pointer.
    parity = None
    if not args.no_parity_check:
        barrier()
        snap_q, snap_p = queue.clone(), ptr.clone()
        n_par = K // (world * B) + 3
        k_save = k_raw.clone()
        keys_log = []
        for s_ in range(n_par):
            k_raw.copy_(torch.roll(k_save, shifts=s_ + 1, dims=1))
            r = step()
            keys_log.append(r["k_hat"].clone())
        barrier()
        got_q, got_p = queue.clone(), ptr.clone()
        queue.copy_(snap_q)
        ptr.copy_(snap_p)
        for kh in keys_log:
            if world > 1:
                dist.all_gather_into_tensor(gathered, kh)
                ops.enqueue_(queue, gathered, ptr)
            else:
                p0 = int(ptr)
                queue[:, p0:p0 + B] = kh.T.to(queue.dtype)
                ptr[0] = (p0 + B) % K
        barrier()
        same_replay = bool(torch.equal(got_q, queue)) and bool(torch.equal(got_p, ptr))
        same_ranks = True
        if world > 1:
            ref_q, ref_p = got_q.clone(), got_p.clone()
            dist.broadcast(ref_q, 0)
            dist.broadcast(ref_p, 0)
            flag = torch.tensor([int(torch.equal(ref_q, got_q) and torch.equal(ref_p, got_p)), int(same_replay)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same_ranks, same_replay = bool(flag[0].item()), bool(flag[1].item())
        k_raw.copy_(k_save)
        parity = {"ok": same_ranks and same_replay, "queue_identical_across_ranks": same_ranks,
                  "matches_replay": same_replay, "steps": n_par, "ptr": int(got_p.item()),
                  "replay": "all_gather_into_tensor + rmcl_enqueue" if world > 1 else "eager torch strided copy (objectives.py:244-248)",
                  "queue_sum64": float(got_q.double().sum().item())}
        del snap_q, got_q, keys_log

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
    #      evicts the 33.5 MB queue from the 126 MB L2 before every InfoNCE pass).  N>1: the exchange is timed as the
    #      step runs it — the kernel the step uses, launched right after the InfoNCE while the peers are at whatever
    #      point of their own step they are, so its duration includes the wait for the slowest rank.
    kern_ms = {"ema": 0.0, "infonce_prep": 0.0, "infonce_partial": 0.0, "infonce_finalize": 0.0, "infonce_call": 0.0, "enqueue": 0.0}
    reps = min(args.steps, 20)
    if world > 1:
        t = torch.tensor([reps], device=dev)
        dist.broadcast(t, 0)
        reps = int(t.item())
    ops.profile_enable(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    for _ in range(reps):
        ev[0].record()
        ops.ema_multi_(plan, m, check_storage=False)   # the bracket must hold the kernel, not the host-side address check
        ev[1].record()
        r = ops.infonce_fwd_bwd(q, k_raw, queue, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        ev[2].record()
        exchange_and_enqueue(r["k_hat"])
        ev[3].record()
        torch.cuda.synchronize()
        st = ops.profile_infonce_ms()
        kern_ms["ema"] += ev[0].elapsed_time(ev[1]) / reps
        kern_ms["infonce_prep"] += st["prep"] / reps
        kern_ms["infonce_partial"] += st["partial"] / reps
        kern_ms["infonce_finalize"] += st["finalize"] / reps
        kern_ms["infonce_call"] += ev[1].elapsed_time(ev[2]) / reps
        kern_ms["enqueue"] += ev[2].elapsed_time(ev[3]) / reps
    ops.profile_enable(False)
    if exchange_info is not None:
        us_alone = exchange_info.get("us_p2p" if p2p is not None else "us_nccl")
        exchange_info["us_in_step"] = kern_ms["enqueue"] * 1e3
        exchange_info["us_skew"] = max(0.0, kern_ms["enqueue"] * 1e3 - us_alone)
        exchange_info["note"] = ("us_<impl>: the exchange alone, all ranks entering together; us_in_step: the same launch inside the "
                                 "step sequence (serialised behind the InfoNCE for this measurement); us_skew = their difference = "
                                 "waiting for the slowest rank.  In the timed step the exchange runs on a side stream under the EMA.")

    flops_infonce = 4.0 * B * C * (K + 1)          # fused fwd (q.K^T) + bwd (P.K): 2 GEMMs of 2*B*C*K
    traffic, traffic_meta = traffic_table()
    kernels = {}
    kernels["ema"] = hbm_entry(12.0 * n_params, kern_ms["ema"], pk_peak)                 # read k, read q, write k (fp32)
    infonce_names = ops.infonce_launch_names(B, C, K, queue.dtype, path)
    fused = infonce_names == ("infonce_fused_kernel",)
    if not fused:       # three-launch chain: the split-K flash kernel bracketed on its own
        kernels["infonce_partial"] = tensor_entry(flops_infonce, kern_ms["infonce_partial"], pk_peak)
    kernels["infonce_call"] = dict(tensor_entry(flops_infonce, kern_ms["infonce_call"], pk_peak), launches=list(infonce_names),
                                   what="the whole rmcl_infonce_fwd_bwd call inside the step" +
                                        (": ONE cooperative kernel (prep rows | tcgen05 flash pass | finalize rows)" if fused else ""))
    kernels["enqueue"] = hbm_entry(world * B * C * (4 + 2), kern_ms["enqueue"], pk_peak)  # read fp32 keys, write bf16 columns
    kernels["enqueue"]["kernel"] = ("gather_enqueue_p2p_kernel" if p2p is not None else
                                    ("ncclAllGather + enqueue_kernel" if world > 1 else "enqueue_kernel"))
    for name in kernels:
        kernels[name]["traffic"] = traffic.get(name)
    # ---- the tcgen05 InfoNCE kernel over CONSECUTIVE launches (one event pair around the whole batch).
    #      An event pair around a single launch also times launch/event latency, so the bracketed number above is an
    #      upper bound.  Here distinct queue copies (together > the 126 MB L2) are cycled so that every launch streams
    #      its queue from HBM as it does in the step.
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

        def call_batch():
            for j in range(n_b2b):
                ops.infonce_fwd_bwd(q, k_raw, queues[j % n_copies], tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))

        if "infonce_partial" not in kernels:    # fused build: the flash pass alone exists only as this measurement
            kernels["infonce_partial"] = {"bound": "tensor", "unit": "TFLOP/s", "traffic": traffic.get("infonce_partial"),
                                          "what": "the tcgen05 flash pass alone (RMCL_INFONCE_DEBUG_PARTIAL_ONLY), consecutive launches"}
        for tag, fn in (("infonce_partial", partial_batch), ("infonce_call", call_batch)):
            fn()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()          # replayed as one graph so that the host cannot starve the stream
            with torch.cuda.graph(graph):
                fn()
            graph.replay()
            ms_b2b = min(timed(graph.replay, 1) for _ in range(5)) / n_b2b
            kp = kernels[tag]
            kp["ms_consecutive"] = ms_b2b
            kp["achieved_consecutive"] = flops_infonce / (ms_b2b * 1e-3) / 1e12
            kp["frac_consecutive"] = kp["achieved_consecutive"] / pk_peak["tf_sust"]
            kp["method"] = ("ms/achieved/frac: one CUDA-event pair around the launch(es) inside the step (includes launch + event "
                            "latency), against the BURST peak; *_consecutive: one event pair around %d back-to-back launches (one "
                            "CUDA graph) cycling over %d queue copies (> L2), divided by the count, against the SUSTAINED peak"
                            % (n_b2b, n_copies))
            del graph
        del queues

    extra = world == 1 and not args.no_pgd
    # ---- cfg3: the PGD update kernel alone (not part of the cfg2 step): B=128, pixel and embedding perturbations,
    #      every mode; 12 B/element (read g, read delta, write delta)
    if extra:
        for tag, shape, mode, lr, eps in (("pgd_pixel_ref_linf", (128, 3, 384, 384), "ref_linf", 0.05, 8 / 255),
                                          ("pgd_pixel_sign_linf", (128, 3, 384, 384), "sign_linf", 2 / 255, 8 / 255),
                                          ("pgd_pixel_l2", (128, 3, 384, 384), "l2", 0.5, 1.0),
                                          ("pgd_embed_ref_linf", (128, 185, 768), "ref_linf", 0.05, 8 / 255),
                                          ("pgd_embed_sign_linf", (128, 185, 768), "sign_linf", 2 / 255, 8 / 255),
                                          ("pgd_embed_l2", (128, 185, 768), "l2", 0.5, 1.0)):
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
            kernels[tag] = dict(hbm_entry(12.0 * grad.numel(), e0.elapsed_time(e1) / 5, pk_peak), traffic=traffic.get(tag),
                                shape=list(shape), timing="5 consecutive launches (the 5 PGD steps of cfg3), one event pair")
            del grad, delta
    # ---- cfg5 per-GPU InfoNCE (B512 C768 K262144 bf16: the two-pass tcgen05 variant, timed as one call;
    #      queue 403 MB + P~ 268 MB > L2, so every call streams from HBM)
    if extra and path in ("auto", "tcgen05"):
        B5, C5, K5 = 512, 768, 262144
        g5 = torch.Generator(device=dev).manual_seed(5)
        q5 = torch.randn(B5, C5, device=dev, generator=g5).bfloat16()
        k5 = torch.randn(B5, C5, device=dev, generator=g5).bfloat16()
        queue5 = torch.nn.functional.normalize(torch.randn(C5, K5, device=dev, generator=g5), dim=0).bfloat16()
        for _ in range(3):
            ops.infonce_fwd_bwd(q5, k5, queue5, tau, normalize_k=True, path=path, want=("loss", "dq", "k_hat"))
        ms5 = timed(lambda: ops.infonce_fwd_bwd(q5, k5, queue5, tau, normalize_k=True, path=path,
                                                want=("loss", "dq", "k_hat")), 10) / 10
        kernels["infonce_cfg5_two_pass"] = dict(tensor_entry(4.0 * B5 * C5 * (K5 + 1), ms5, pk_peak),
                                                traffic=traffic.get("infonce_cfg5_two_pass"), shape=[B5, C5, K5],
                                                timing="10 whole calls, one event pair")
        del q5, k5, queue5
    # ---- cfg4 per-GPU InfoNCE shape (B128 C128 K65536): arithmetic intensity 2*B = 256 flop per queue byte puts it
    #      at the HBM/tensor ridge; whole call by CUDA-graph replay over queue copies larger than L2, so every call
    #      streams its queue from HBM.  bf16 queue (main step under autocast) and fp32 queue (the PGD inner loss,
    #      pgd_attack_vilt.py:141: the fp32-accurate split-operand tcgen05 path)
    if extra and path in ("auto", "tcgen05"):
        B4, C4, K4 = 128, 128, 65536
        g4 = torch.Generator(device=dev).manual_seed(4)
        q4 = torch.randn(B4, C4, device=dev, generator=g4)
        k4 = torch.randn(B4, C4, device=dev, generator=g4)
        for tag, qdt, ncopy in (("infonce_cfg4_b128_c128_call", torch.bfloat16, 12), ("infonce_cfg4_b128_c128_fp32_call", torch.float32, 6)):
            queues4 = [torch.nn.functional.normalize(torch.randn(C4, K4, device=dev, generator=g4), dim=0).to(qdt) for _ in range(ncopy)]
            qq4, kk4 = (q4.bfloat16(), k4.bfloat16()) if qdt == torch.bfloat16 else (q4, k4)
            try:
                for qu in queues4[:2]:
                    ops.infonce_fwd_bwd(qq4, kk4, qu, tau, normalize_k=True, path="auto", want=("loss", "dq", "k_hat"))
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for j in range(4 * ncopy):
                        ops.infonce_fwd_bwd(qq4, kk4, queues4[j % ncopy], tau, normalize_k=True, path="auto", want=("loss", "dq", "k_hat"))
                graph.replay()
                ms4 = min(timed(graph.replay, 1) for _ in range(5)) / (4 * ncopy)
                esz = 2 if qdt == torch.bfloat16 else 4
                kernels[tag] = dict(hbm_entry(C4 * K4 * esz, ms4, pk_peak), tflops=4.0 * B4 * C4 * (K4 + 1) / (ms4 * 1e-3) / 1e12,
                                    traffic=None, shape=[B4, C4, K4],
                                    timing="CUDA graph of %d whole calls over %d queue copies (> L2)" % (4 * ncopy, ncopy))
                del graph
            except Exception as e:  # noqa: BLE001
                kernels[tag] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            del queues4
    # ---- fp32 queue at cfg2 (the reference's buffer dtype and its PGD-inner precision): fp32-accurate InfoNCE
    if extra and path in ("auto", "tcgen05"):
        q32, k32 = q.float(), k_raw.float()
        queues32 = [torch.randn(C, K, device=dev, generator=g) for _ in range(3)]
        try:
            for qu in queues32:
                ops.infonce_fwd_bwd(q32, k32, qu, tau, normalize_k=True, path="auto", want=("loss", "dq", "k_hat"))
            ms32 = timed(lambda: [ops.infonce_fwd_bwd(q32, k32, qu, tau, normalize_k=True, path="auto", want=("loss", "dq", "k_hat"))
                                  for qu in queues32], 5) / 15
            kernels["infonce_cfg2_fp32_call"] = dict(tensor_entry(flops_infonce, ms32, pk_peak), shape=[B, C, K], traffic=None,
                                                     what="fp32 queue + fp32 q/k, fp32-accurate path chosen by auto dispatch; flops "
                                                          "counted once (the split-operand path spends 3x per GEMM)",
                                                     timing="15 whole calls over 3 queue copies (201 MB > L2), one event pair")
        except Exception as e:  # noqa: BLE001
            kernels["infonce_cfg2_fp32_call"] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        del queues32, q32, k32
    # ---- Barlow-Twins objective at the reference's size (batch 128, projector 8192) and at an 8-GPU gathered batch (1024)
    if extra:
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
                ent = tensor_entry((6.0 * Bb * Bb * Db) if bpath == "gram" else (4.0 * Bb * Db * Db), msb, pk_peak)
            else:
                ent = hbm_entry(3.0 * Bb * Db * 4, msb, pk_peak)         # read q, k (fp32), write dq
            kernels[tag] = dict(ent, traffic=traffic.get(tag), shape=[Bb, Db], timing="CUDA graph of 10 calls")
            del graph, qb_, kb_
    for name in ("infonce_prep", "infonce_finalize"):
        if kern_ms[name] > 0 and not fused:
            kernels[name] = {"ms": kern_ms[name]}
    dominant = max(("ema", "infonce_call", "enqueue"), key=lambda n: kern_ms[n])
    roofline = dict(kernels[dominant], kernel={"ema": "ema_multi_kernel", "infonce_call": "+".join(infonce_names),
                                               "enqueue": kernels["enqueue"]["kernel"]}[dominant],
                    peak_source=pk_peak["source"], traffic_capture=traffic_meta)

    launches = ["ema_multi_kernel"] + list(infonce_names) + \
               ["gather_enqueue_p2p_kernel" if p2p is not None else "enqueue_kernel"]
    n_blocks = len(blocks)
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, **CFG, "arithmetic": "InfoNCE: bf16 queue/q/k operands, fp32 accumulation and statistics; EMA: fp32 (bit-exact with ATen); enqueue: fp32 keys -> bf16 queue", "per_gpu_batch": B, "global_batch": world * B, "infonce_path": path,
                   "parallelism": f"dp{world}", "exchange": exchange_info,
                   "streams": "single stream" if world == 1 else "key exchange + enqueue on a high-priority side stream under the EMA, joined in front of the next InfoNCE", "unit_of_value": "256-sample rank-steps per second, summed over ranks",
                   "l2": "inputs larger than L2: each step streams 1.34 GB of parameters (EMA) between InfoNCE passes; L2 is 126 MB",
                   "timing": f"{n_blocks} blocks of exactly {args.steps} steps, each bracketed by barrier+synchronize, CUDA events, max over ranks; "
                             f"ms_per_step = median block / {args.steps}"},
        "timed_blocks": {"n": n_blocks, "ms_min": min(blocks), "ms_median": total_ms, "ms_max": max(blocks), "timed_s": sum(blocks) * 1e-3},
        "roofline": roofline, "kernels": kernels,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "api": e2e_api, "blocks": len(e2e_blocks)},
        "gpu_launches": len(launches) * args.steps, "launches_per_step": launches,
        "parity_check": parity,
        "clocks": clocks, "loss": loss_val,
    }
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        line["gpu_baseline"] = gpu_baseline(dev, shapes, q, k_raw, queue, args.steps, use_compile=not args.no_compile)
        line["gpu_baseline"]["speedup_vs_eager"] = value / line["gpu_baseline"]["eager"]["value"]
        tc = line["gpu_baseline"].get("torch_compile", {})
        if "value" in tc:
            line["gpu_baseline"]["speedup_vs_torch_compile"] = value / tc["value"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample = time_cpu(steps=5, warmup=2)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cfg4(args):
    """BASELINE configs[3]: the full RMCL step (tools/full_step.py) reported in the bench line format."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import full_step
    full_step.bench_main(args, METRIC, ClockSampler)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg4"],
                    help="cfg2 (default, BASELINE configs[1]): kernels-only step; cfg4 (configs[3]): full RMCL step around a ViLT-B/32-shaped torch backbone")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1 key exchange: fused peer-memory kernel (p2p) or ncclAllGather + enqueue (nccl); auto = p2p if available")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-compile", action="store_true", help="gpu_baseline: eager only (skip the torch.compile arm)")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-pgd", action="store_true", help="skip the cfg3 PGD-kernel, cfg4/cfg5 InfoNCE and Barlow roofline lines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "cfg4":
        run_cfg4(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()

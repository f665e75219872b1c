#!/usr/bin/env python
"""bench.py — headline benchmark of the latent-space clustering hot path on B200.

Workload at every N (weak scaling, per-GPU work fixed) = BASELINE.json configs[1]:
  "DEC ClusteringLayer forward/backward + target distribution, N=1M latent points
   d=9 K=8 alpha=1"  — one STEP is the reference chain over one latent set, every N-sized
  result the reference keeps (q, labels, p, dL/dz) written to HBM:
     dec_assign          (q, labels, f; np.round(q,5))                 networks.py:279-288, models.py:92-94
     dec_target_kl_grad  (p = target_distribution(q) written out, loss,
                          dL/dz, dL/dmu; q recomputed in registers)    models.py:1320-1322, 1124-1127
  On one GPU both passes run as ONE cooperative kernel (`ops.dec_step`: a grid-wide barrier all-reduces f
  between them); with N>1 they are two kernels with the two packed-statistics exchanges between/after
  them.  `--two-kernel` / `--unfused` run the two- / three-kernel chains instead (also timed under "extra").

Contract (one JSON line on rank 0): metric/value/unit, n_gpus, steps, warmup,
ms_per_step, higher_is_better, scaling, vs_baseline, dtype, data, config, clocks,
e2e, gpu_launches, roofline, cpu_baseline (+ "extra": GMM EM iteration and the
fused d=32 latent-buffer DEC pass).  `--impl reference` times the reference's own
CPU op chain (oracle port, torch CPU autograd + numpy, float64 as the reference
runs it) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "latent points/sec per DEC step (ClusteringLayer fwd+bwd + target distribution)"
UNIT = "points/s"
N_PER_GPU, D, K, ALPHA, GAMMA = 1_000_000, 9, 8, 1.0, 1e-3
N_SETS = 4                      # distinct input/output sets cycled so no step re-finds its data in L2
GRAPH_STEPS = 20                # consecutive steps captured into one CUDA graph (a multiple of N_SETS; divides the
                                # driver's --steps 20, so a timed region is whole multi-step graph launches)
WORKLOAD = "DEC fwd/bwd + target distribution, N=1M latent points per GPU, d=9, K=8, alpha=1 (BASELINE configs[1])"




def workload_config(n_gpus: int) -> dict:
    """The `config` object both arms print verbatim (same keys, same values) — what differs between the arms
    (launch mode, parallelism, host) lives under `details`."""
    return {"workload": WORKLOAD, "n_points_per_gpu": N_PER_GPU, "n_points_total": N_PER_GPU * n_gpus, "d": D, "K": K,
            "alpha": ALPHA, "gamma": GAMMA, "round_decimals": 5,
            "l2": f"GPU arm: inputs/outputs rotate over {N_SETS} sets ({N_SETS * 140} MB) > 126 MB L2, so no step "
                  "re-finds its data in L2; reference arm: host memory"}


_JSON_FD = None


def protect_stdout():
    """Keep stdout for the ONE JSON line: native libraries (the NCCL version banner) write to fd 1 too,
    so fd 1 is pointed at stderr and the JSON line goes out through a saved duplicate of the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def dbg(msg):
    if os.environ.get("SCC_BENCH_DEBUG"):
        print(f"[bench r{os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Polls NVML (or nvidia-smi) for SM clock and throttle reasons while the GPU is under load."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _poll_once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._poll_once()
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------ reference arm
def reference_step_fn(n, dtype_name="float64"):
    """The reference's own CPU op chain for one step (oracle port: same torch/numpy calls)."""
    import torch
    from oracle import dec as odec
    from spectrogram_cube_clustering_b200 import synth
    z, mu = synth.latent_points(n, D, K, rank=0, device="cpu")
    dt = torch.float64 if dtype_name == "float64" else torch.float32
    z, mu = z.to(dt), mu.to(dt)
    return lambda: odec.torch_dec_step(z, mu, ALPHA, GAMMA)


def time_reference(steps, warmup, budget_s=150.0):
    import torch
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    n = N_PER_GPU
    fn = reference_step_fn(n)
    t0 = time.perf_counter(); fn(); first = time.perf_counter() - t0
    if first * (steps + warmup) > budget_s:            # bound the sample so the run ends in minutes
        n = max(10_000, int(n * budget_s / (first * (steps + warmup))))
        fn = reference_step_fn(n)
    for _ in range(max(0, warmup - 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=n / dt, ms_per_step=dt * 1e3, cores=cores, n=n,
                sample=f"{steps} steps of the float64 reference op chain (torch CPU autograd + numpy "
                       f"target_distribution) on {n} of {N_PER_GPU} points, {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "details": {"host": "reference CPU path (PyTorch CPU autograd + numpy), no GPU; float64 as models.py:965 runs it",
                    "n_points_sampled": r["n"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from spectrogram_cube_clustering_b200 import ops, synth
    from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from spectrogram_cube_clustering_b200.latent_buffer import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(dev) if world > 1 else {"bound": False}     # pinned staging on the GPU-local node
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    n_total = N_PER_GPU * world
    hbm_peak, peak_src = peaks()

    _, mu = synth.latent_points(16, D, K, device=dev)
    sets = []
    for s in range(N_SETS):
        z, _ = synth.latent_points(N_PER_GPU, D, K, rank=rank * N_SETS + s, device=dev)
        sets.append(dict(z=z, q=torch.empty(N_PER_GPU, K, device=dev), p=torch.empty(N_PER_GPU, K, device=dev),
                         dz=torch.empty(N_PER_GPU, D, device=dev),
                         labels=torch.empty(N_PER_GPU, dtype=torch.int32, device=dev),
                         st1=torch.empty(K + 1, dtype=torch.float64, device=dev),
                         st2=torch.empty(K * D + 2, dtype=torch.float64, device=dev)))
    scale = GAMMA / n_total

    exchange = None
    if world > 1 and not args.nccl:
        try:
            from spectrogram_cube_clustering_b200.latent_buffer import PeerExchange
            exchange = PeerExchange(group, dev, 1 + 16 + 16 * D + 16 * (D * (D + 1) // 2))
            chk = torch.arange(K * D + 2, dtype=torch.float64, device=dev) * (rank + 1)
            ref = chk.clone(); dist.all_reduce(ref, group=group)
            exchange.all_reduce(chk)
            torch.cuda.synchronize()
            assert torch.equal(chk, ref), "peer exchange disagrees with NCCL"
            dbg("peer exchange verified against NCCL")
        except Exception as exc:
            print(f"[bench] peer-memory exchange unavailable ({exc!r}); using NCCL all_reduce", file=sys.stderr)
            exchange = None

    def allreduce(t):
        if world > 1:
            if exchange is not None:
                exchange.all_reduce(t)
            else:
                dist.all_reduce(t, group=group)

    def k_assign(s):
        ops.dec_assign(s["z"], mu, ALPHA, 5, out_q=s["q"], out_labels=s["labels"], out_stats=s["st1"])

    def k_target(s):
        ops.dec_target(s["q"], s["st1"], 5, out=s["p"])

    def k_grad(s):
        ops.dec_kl_grad(s["z"], mu, ALPHA, p=s["p"], scale=scale, out_dz=s["dz"], out_stats=s["st2"])

    def k_tgrad(s):         # target distribution + KL loss + gradients in one pass; p still materialised
        ops.dec_target_kl_grad(s["z"], mu, s["st1"], ALPHA, 5, scale, out_p=s["p"], out_dz=s["dz"], out_stats=s["st2"])

    fused_ex = exchange is not None and args.fused_exchange      # measured ~2 % slower than the stand-alone kernel
    unfused = args.unfused
    # one cooperative kernel per step; with N > 1 the all-reduce of f runs inside it over NVLink peer memory
    one_kernel = (world == 1 or exchange is not None) and not (unfused or args.two_kernel or fused_ex)
    ex_desc = exchange.desc if (exchange is not None and world > 1) else None

    def k_step(s):          # assign pass + grid-wide all-reduce of f + target/gradient pass in one cooperative kernel
        ops.dec_step(s["z"], mu, ALPHA, 5, scale, out_q=s["q"], out_labels=s["labels"], out_p=s["p"], out_dz=s["dz"],
                     out_f=s["st1"], out_stats=s["st2"], exchange=ex_desc)

    def step_unfused(s):
        k_assign(s); allreduce(s["st1"]); k_target(s); k_grad(s); allreduce(s["st2"])

    def step(s):
        if one_kernel:
            k_step(s)
        elif unfused:
            step_unfused(s)
        elif fused_ex:      # collectives ride on the kernels: push in the producers' tails, pull in the consumers
            ex = exchange.desc
            ops.dec_assign(s["z"], mu, ALPHA, 5, out_q=s["q"], out_labels=s["labels"], out_stats=s["st1"], push=ex)
            ops.dec_target_kl_grad(s["z"], mu, None, ALPHA, 5, scale, out_p=s["p"], out_dz=s["dz"],
                                   out_stats=s["st2"], pull_f=ex, push=ex)
        else:
            k_assign(s); allreduce(s["st1"]); k_tgrad(s); allreduce(s["st2"])

    dbg('inputs ready')
    if fused_ex:            # one fused step must reproduce the NCCL-reduced statistics
        s0 = sets[0]
        k_assign(s0); f_ref = s0["st1"].clone(); dist.all_reduce(f_ref, group=group)
        ops.dec_target_kl_grad(s0["z"], mu, f_ref, ALPHA, 5, scale, out_p=s0["p"], out_dz=s0["dz"], out_stats=s0["st2"])
        g_ref = s0["st2"].clone(); dist.all_reduce(g_ref, group=group)
        step(s0)
        torch.cuda.synchronize()
        assert torch.allclose(s0["st2"], g_ref, rtol=1e-12, atol=0), "fused exchange: gradient statistics differ from NCCL"
        dbg("fused exchange verified against NCCL")
    if one_kernel and world > 1:    # the in-kernel exchange must reproduce the NCCL-reduced two-kernel chain
        s0 = sets[0]
        k_assign(s0); f_ref = s0["st1"].clone(); dist.all_reduce(f_ref, group=group)
        ops.dec_target_kl_grad(s0["z"], mu, f_ref, ALPHA, 5, scale, out_p=s0["p"], out_dz=s0["dz"], out_stats=s0["st2"])
        g_ref = s0["st2"].clone(); dist.all_reduce(g_ref, group=group)
        p_ref = s0["p"].clone()
        k_step(s0)
        torch.cuda.synchronize()
        # (f: per-thread fp32 partial sums are grouped differently in the two kernels -> ~1e-7 relative)
        f_err = ((s0["st1"] - f_ref).abs() / f_ref.abs().clamp_min(1e-300)).max().item()
        assert f_err <= 1e-6, f"one-kernel step: column sums differ from NCCL (max rel {f_err:.3e})"
        assert torch.allclose(s0["st2"], g_ref, rtol=1e-4, atol=1e-18), "one-kernel step: gradient statistics differ from NCCL"
        assert (s0["p"] - p_ref).abs().max().item() <= 1.01e-5, "one-kernel step: target distribution differs"
        dbg("one-kernel step verified against NCCL")
    # warm-up (eager): also creates workspaces and primes NCCL
    for w in range(max(args.warmup, 3)):
        step(sets[w % N_SETS])
    torch.cuda.synchronize()

    dbg('eager warm-up done')
    # CUDA graphs: one per input set (pointers are baked in)
    graphs, graph_all, use_graphs = [], None, not args.no_graphs
    if use_graphs:
        try:
            cap_stream = torch.cuda.Stream()
            with torch.cuda.stream(cap_stream):
                for s in sets:
                    step(s)                      # workspaces for the capture stream
                cap_stream.synchronize()
                for s in sets:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cap_stream):
                        step(s)
                    graphs.append(g)
                # GRAPH_STEPS consecutive steps in ONE graph: the programmatic (PDL) edges between kernels then
                # also span step boundaries and there is one graph launch per GRAPH_STEPS steps
                graph_all = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph_all, stream=cap_stream):
                    for _ in range(GRAPH_STEPS // N_SETS):
                        for s in sets:
                            step(s)
            torch.cuda.synchronize()
            for g in graphs:
                g.replay()
            graph_all.replay()
            torch.cuda.synchronize()
        except Exception as exc:                  # pragma: no cover
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({exc}); timing eager launches", file=sys.stderr)
            graphs, use_graphs = [], False

    def run_step(i):
        if use_graphs:
            graphs[i % N_SETS].replay()
        else:
            step(sets[i % N_SETS])

    def run_steps(n_steps):
        """Exactly n_steps steps: whole GRAPH_STEPS-step graphs, then single-step graphs for the remainder."""
        i = 0
        if use_graphs and not args.single_step_graphs:
            while i + GRAPH_STEPS <= n_steps:
                graph_all.replay()
                i += GRAPH_STEPS
        while i < n_steps:
            run_step(i)
            i += 1

    dbg(f'graphs={use_graphs}')
    # untimed: the W warm-up steps through the timed launch path, then ~0.25 s of the same load so that
    # the timed region starts with the GPU at its steady clocks (a fixed, rank-uniform step count)
    run_steps(max(args.warmup, 3))
    for _ in range(0, 4096, 512):
        run_steps(512)
        torch.cuda.synchronize()
    sampler = ClockSampler(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dbg('timed region')
    # ---------------- timed region: EXACTLY `steps` steps ----------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    if world > 1:
        # pre-roll (untimed, same stream, no host sync after it): the ranks leave the host barrier tens of
        # microseconds apart, and with K = 20 steps (~1 ms) that launch skew would be charged to the first timed
        # step's exchange.  The in-kernel exchanges of these steps align the ranks ON THE DEVICE; the events
        # below are enqueued while they still run, so they bracket exactly K steps of aligned ranks.
        run_steps(2 * GRAPH_STEPS)
    ev0.record()
    run_steps(args.steps)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    # keep the same load running a little longer so NVML gets samples even for very short regions
    # (a FIXED step count derived from the max-reduced time: every rank must issue the same collectives)
    n_extra = max(8, min(20000, int(150.0 / max(ms_total / args.steps, 1e-3))))
    for i in range(0, n_extra, 64):
        run_steps(64)
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    sampler.stop()
    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step * 1e-3)

    dbg('per-kernel pass')
    # ---------------- per-kernel durations ----------------
    # One CUDA graph per kernel holding GRAPH_STEPS launches that rotate over the N_SETS input sets; the
    # replays are timed with CUDA events on the launching stream, so a kernel's figure is its average
    # launch duration in steady state (in-graph launch gaps included), inputs cold in L2 like the step.
    KERNEL_GRAPH_ROUNDS = GRAPH_STEPS // N_SETS     # as many launches per replay as the step graphs hold

    def graph_of(fn):
        cs = torch.cuda.Stream()
        with torch.cuda.stream(cs):
            for s_ in sets:
                fn(s_)
            cs.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cs):
                for _ in range(KERNEL_GRAPH_ROUNDS):
                    for s_ in sets:
                        fn(s_)
        torch.cuda.synchronize()
        return g

    def time_graph(g, launches, reps):
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * launches)

    reps = max(5, min(args.steps, 50))
    kfns = {"dec_step": k_step} if one_kernel else (
        {"dec_assign": k_assign, "dec_target": k_target, "dec_kl_grad": k_grad} if unfused else
        {"dec_assign": k_assign, "dec_target_kl_grad": k_tgrad})
    for s_ in sets:                      # q / f / p of every set valid for the stand-alone kernels
        k_assign(s_); k_target(s_)
    kavg = {k: time_graph(graph_of(fn), N_SETS * KERNEL_GRAPH_ROUNDS, max(3, reps // 2)) for k, fn in kfns.items()}
    alg_bytes = {"dec_assign": 4 * D + 4 * K + 4, "dec_target": 8 * K, "dec_kl_grad": 8 * D + 4 * K,
                 "dec_target_kl_grad": 8 * D + 4 * K,
                 "dec_step": (4 * D + 4 * K + 4) + (8 * D + 4 * K)}                    # per point
    dominant = max(kavg, key=kavg.get)
    achieved = alg_bytes[dominant] * N_PER_GPU / (kavg[dominant] * 1e-3) / 1e9
    step_bytes = sum(alg_bytes[k] for k in kavg)
    traffic = None
    try:          # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        traffic = tr.get(dominant)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_note": "ncu --set full cold-cache capture (profiles/traffic.json, profiles/r02_ncu_kernels.txt); rows "
                                "written by the kernel are still in the 126 MB L2 when it ends, so DRAM writes are "
                                "below the algorithmic store bytes",
                "peak_source": peak_src,
                "algorithmic_bytes_per_point": alg_bytes[dominant],
                "algorithmic_bytes": {"dec_target_kl_grad": "z read 4d + p written 4K + dz written 4d (q is recomputed in registers)",
                                      "dec_step": "pass 1: z read 4d + q written 4K + labels 4; pass 2: z read 4d (mostly from "
                                                  "L2) + p written 4K + dz written 4d"}.get(dominant),
                "kernels_ms": kavg,
                "kernels_gbs": {k: alg_bytes[k] * N_PER_GPU / (kavg[k] * 1e-3) / 1e9 for k in kavg},
                "step_bytes_per_point": step_bytes,
                "step_gbs": step_bytes * N_PER_GPU / (ms_per_step * 1e-3) / 1e9,
                "step_frac": step_bytes * N_PER_GPU / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                "fp32_note": "the kernel is FP32-issue/latency-bound, not HBM-bound (SURVEY.md 8d; profiles/)"}

    unfused_extra, two_extra = None, None
    if world == 1 and not unfused:       # the three-kernel chain of the earlier sessions, for comparison
        g3 = graph_of(lambda s_: (k_assign(s_), k_target(s_), k_grad(s_)))
        ms3 = time_graph(g3, N_SETS * KERNEL_GRAPH_ROUNDS, max(3, reps // 2))
        unfused_extra = {"workload": "dec_assign -> dec_target -> dec_kl_grad(p): the 3-kernel chain (240 B/point)",
                         "ms": ms3, "points_per_s": N_PER_GPU / (ms3 * 1e-3),
                         "hbm_frac": 240 * N_PER_GPU / (ms3 * 1e-3) / 1e9 / hbm_peak}
    if one_kernel:                       # the two-kernel chain every rank runs when N > 1
        g2 = graph_of(lambda s_: (k_assign(s_), k_tgrad(s_)))
        ms2 = time_graph(g2, N_SETS * KERNEL_GRAPH_ROUNDS, max(3, reps // 2))
        two_extra = {"workload": "dec_assign -> dec_target_kl_grad: the 2-kernel chain (176 B/point)",
                     "ms": ms2, "points_per_s": N_PER_GPU / (ms2 * 1e-3),
                     "hbm_frac": 176 * N_PER_GPU / (ms2 * 1e-3) / 1e9 / hbm_peak}

    dbg('e2e pass')
    # ---------------- end to end: host buffers in, host results out, every step ----------------
    zh = [s["z"].cpu().pin_memory() for s in sets[:2]]
    mu_h = mu.cpu().pin_memory()
    res_h = torch.empty(K * D + 2 + K + 1, dtype=torch.float64).pin_memory()
    zd2 = [torch.empty(N_PER_GPU, D, device=dev) for _ in range(2)]
    mud = torch.empty_like(mu)
    sd = sets[0]
    copy_stream = torch.cuda.Stream()
    up_done = [torch.cuda.Event() for _ in range(2)]
    buf_free = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream()
    for ev in buf_free:
        ev.record(main)

    def upload(i):
        """H2D of step i's latent set on the copy stream (double-buffered: overlaps step i-1's kernels)."""
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(buf_free[b])
            zd2[b].copy_(zh[b], non_blocking=True)
            up_done[b].record(copy_stream)

    def e2e_step(i):
        b = i % 2
        upload(i + 1)                                       # next step's input is already on its way
        main.wait_event(up_done[b])
        zd = zd2[b]
        mud.copy_(mu_h, non_blocking=True)
        if one_kernel:
            ops.dec_step(zd, mud, ALPHA, 5, scale, out_q=sd["q"], out_labels=sd["labels"], out_p=sd["p"], out_dz=sd["dz"],
                         out_f=sd["st1"], out_stats=sd["st2"], exchange=ex_desc)
        elif unfused:
            ops.dec_assign(zd, mud, ALPHA, 5, out_q=sd["q"], out_labels=sd["labels"], out_stats=sd["st1"])
            allreduce(sd["st1"])
            ops.dec_target(sd["q"], sd["st1"], 5, out=sd["p"])
            ops.dec_kl_grad(zd, mud, ALPHA, p=sd["p"], scale=scale, out_dz=sd["dz"], out_stats=sd["st2"])
            allreduce(sd["st2"])
        elif fused_ex:
            ex = exchange.desc
            ops.dec_assign(zd, mud, ALPHA, 5, out_q=sd["q"], out_labels=sd["labels"], out_stats=sd["st1"], push=ex)
            ops.dec_target_kl_grad(zd, mud, None, ALPHA, 5, scale, out_p=sd["p"], out_dz=sd["dz"], out_stats=sd["st2"],
                                   pull_f=ex, push=ex)
        else:
            ops.dec_assign(zd, mud, ALPHA, 5, out_q=sd["q"], out_labels=sd["labels"], out_stats=sd["st1"])
            allreduce(sd["st1"])
            ops.dec_target_kl_grad(zd, mud, sd["st1"], ALPHA, 5, scale, out_p=sd["p"], out_dz=sd["dz"],
                                   out_stats=sd["st2"])
            allreduce(sd["st2"])
        buf_free[b].record(main)
        res_h[:K * D + 2].copy_(sd["st2"], non_blocking=True)
        res_h[K * D + 2:].copy_(sd["st1"], non_blocking=True)
        main.synchronize()                                  # the caller reads loss / dmu on the host
        return float(res_h[0])

    upload(0)
    for i in range(3):
        e2e_step(i)
    e2e_steps = max(5, min(args.steps, 50))
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for i in range(e2e_steps):
        loss_h = e2e_step(i)
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e = {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(zd2[0].numel() * 4 + mud.numel() * 4), "d2h_bytes_per_step": int(res_h.numel() * 8),
           "ms_per_step": e2e_ms, "api": ("ops.dec_step" if one_kernel else "ops.dec_assign + ops.dec_target_kl_grad") + " on a pinned host latent set "
                                         "(upload of step i+1 double-buffered behind step i's kernels); "
                                         "loss, dmu, f, label-change count (664 B) read back every step; the N-sized results "
                                         "(q, labels, p, dz) stay on the device for the next kernels — the reference's "
                                         "batch_eval hands q, labels, z to the host, which this contract does not time",
           "loss": loss_h}

    dbg('e2e done')
    extra = {}
    if not args.no_extra:
        extra = extra_benchmarks(torch, ops, synth, dev, hbm_peak, world, rank, group, exchange)
        if rank == 0 and world == 1 and not args.no_cpu:
            extra.update(gmm_cpu_baselines(torch, synth, extra))
    if unfused_extra is not None:
        extra["dec_step_3_kernels"] = unfused_extra
    if two_extra is not None:
        extra["dec_step_2_kernels"] = two_extra

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = time_reference(3, 1, budget_s=25.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if world == 1:
        parallelism = "single GPU"
    else:
        if one_kernel:
            how = ("NVLink peer-memory exchange INSIDE the one-kernel step, flag-in-data (f: pushed/pulled by the last "
                   "CTA at the grid barrier between the two passes; gradient statistics: pushed/pulled by the kernel's "
                   "last CTA in its tail) — one launch per step and GPU")
        elif fused_ex:
            how = ("NVLink peer-memory exchange fused into the kernels (push in the producer's last CTA, pull in the "
                   "consumer's prologue)")
        elif exchange is not None:
            how = "one-shot NVLink peer-memory exchange kernel"
        else:
            how = "NCCL"
        parallelism = f"latent points sharded over {world} GPU(s); packed f64 stat all-reduce x2/step via " + how
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "details": {"parallelism": parallelism, "numa": numa,
                       "launch": (("one CUDA graph replay per step" if args.single_step_graphs else
                                   f"CUDA graph replays of {GRAPH_STEPS} consecutive steps (rotating over the {N_SETS} input sets), "
                                   "single-step graphs for the remainder")
                                  if use_graphs else "eager launches"),
                       "kernels_per_step": ["dec_step (one cooperative kernel: assign pass, grid-wide all-reduce of f, "
                                            "target + KL-gradient pass)"] if one_kernel else (
                                           ["dec_assign", "dec_target", "dec_kl_grad"] if unfused else
                                           ["dec_assign", "dec_target_kl_grad"]),
                       "timing": "CUDA events around the K steps, max over ranks (multi-GPU: preceded on the same stream by an "
                                 "untimed pre-roll of 40 steps so that the ranks are aligned on the device when the first "
                                 "event is reached); per-kernel durations from "
                                 "CUDA events around replays of single-kernel graphs over the same rotating sets"},
            "clocks": sampler.summary(), "e2e": e2e,
            "gpu_launches": ((1 if one_kernel else (3 if unfused else 2)) +
                             (((0 if (fused_ex or one_kernel) else 2) if exchange is not None else 0)
                              if world > 1 else 0)) * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu, "extra": extra,
        }
        emit(line)
    sys.stdout.flush()
    if world > 1:
        # CUDA graphs that captured NCCL kernels must die before the communicator does; a watchdog
        # guarantees the process exits even if communicator teardown stalls.
        threading.Timer(20.0, lambda: os._exit(0)).start()
        graphs.clear()
        graph_all = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def fp32_peak_tflops():
    """148 SMs x 128 FP32 lanes x 2 FLOP x max SM clock: 74.5 TFLOP/s at 1965 MHz."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mhz = float(json.load(f)["sm_max_mhz"])
    except Exception:
        mhz = 1965.0
    return 148 * 128 * 2 * mhz * 1e6 / 1e12


def extra_benchmarks(torch, ops, synth, dev, hbm_peak, world=1, rank=0, group=None, exchange=None):
    """Secondary figures (not the headline) — the other BASELINE.json configs, at EVERY world size:
      configs[2]  GMM EM fit, 100 iterations, N = 10M TOTAL (strong scaling: N/world points per GPU), d=9 K=16
      configs[3]  DEC refinement step on the 12.5M-point shard of a GPU (weak scaling: 100M points over 8 GPUs), d=32 K=16
      configs[4]  DEC_training epoch on synthetic (N,1,4,101) spectrograms, 131072 per GPU
    Each entry: whole-job points/s (max over ranks of the CUDA-event time), achieved HBM fraction and — for the
    FP32-bound kernels — achieved fraction of the FP32 CUDA-core peak."""
    import torch.distributed as dist
    from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer
    out = {}
    fp32_peak = fp32_peak_tflops()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timeit(fn, reps, flush=None):
        fn(); fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(reps):
            if flush is not None:
                flush.zero_()
            barrier()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += max_over_ranks(e0.elapsed_time(e1))
        return tot / reps

    def shard(n_total):
        lo = rank * (n_total // world)
        return n_total // world if rank < world - 1 else n_total - lo

    try:
        # ---- configs[2] shape: one fused EM iteration (statistics kernel + exchange + device finalize)
        n_total, d, k = 10_000_000, 9, 16
        n = shard(n_total)
        z, _ = synth.latent_points(n, d, k, rank=77 + rank, device=dev)
        buf = LatentBuffer(z, n_total=n_total, group=group, exchange=exchange)
        w0, mu0, cov0 = synth.gmm_initial_state(d, k, dev)
        params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
        means, weights, cov = mu0.clone(), w0.clone(), cov0.clone()
        stats = torch.empty(ops.gmm_stat_doubles(k, d), dtype=torch.float64, device=dev)

        def em():
            buf.gmm_em_pass(k, params, stats, ctrl=ctrl)
            ops.gmm_finalize(stats, n_total, means, weights, cov, pchol, params, ctrl, tol=0.0)
        for _ in range(5):
            em()
        ms = timeit(em, 10)
        flops = 2.0 * k * (d * d + 4 * d)
        tf = flops * n_total / (ms * 1e-3) / 1e12
        out["gmm_em_iteration"] = {"workload": f"fused E+M pass + exchange + device finalize, N=10M total ({n} per GPU), d=9 K=16 "
                                               "(configs[2] shape, strong scaling), after 5 iterations",
                                   "points_per_s": n_total / (ms * 1e-3), "ms": ms,
                                   "hbm_gbs": 4 * d * n_total / (ms * 1e-3) / 1e9,
                                   "hbm_frac": 4 * d * n_total / (ms * 1e-3) / 1e9 / (hbm_peak * world),
                                   "fp32_tflops_algorithmic": tf, "fp32_frac": tf / (fp32_peak * world),
                                   "bound": "fp32 FMA issue (SURVEY.md 8d: the FMA ceiling is 19.9 G points/s per GPU = 11 % "
                                            "of HBM), not HBM"}
        # ---- configs[2]: the whole fit through the scikit-learn-style front end (graph-captured EM iteration,
        # convergence decided on the device, host polls every 25 iterations)
        from spectrogram_cube_clustering_b200.models import GaussianMixture, DEC_training
        from spectrogram_cube_clustering_b200.networks import DEC
        import warnings
        w0c, mu0c, cov0c = synth.gmm_initial_state(d, k, "cpu")
        gm = GaussianMixture(k, max_iter=100, tol=0.0, weights_init=w0c.numpy(), means_init=mu0c.numpy(),
                             covariances_init=cov0c.numpy(), poll_interval=25, group=group)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gm.fit(buf)                                    # warm-up fit (graph capture, workspaces)
            barrier()
            t0 = time.perf_counter()
            gm.fit(buf)
            barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        tf = flops * n_total * gm.n_iter_ / dt / 1e12
        out["gmm_fit_100_iters"] = {"workload": f"GaussianMixture.fit, 100 EM iterations (tol=0), N=10M total ({n} per GPU), d=9 K=16 "
                                                "(configs[2]; strong scaling over the GPUs)",
                                    "seconds": dt, "points_per_s": n_total * gm.n_iter_ / dt, "ms": dt * 1e3 / max(gm.n_iter_, 1),
                                    "n_iter": gm.n_iter_, "lower_bound": gm.lower_bound_, "graph": True,
                                    "hbm_frac": 4 * d * n_total * gm.n_iter_ / dt / 1e9 / (hbm_peak * world),
                                    "fp32_tflops_algorithmic": tf, "fp32_frac": tf / (fp32_peak * world)}
        del z, buf, gm
        # ---- configs[3]: fused latent-buffer DEC step on the 12.5M-point shard of every GPU (weak scaling)
        for name, (d, k) in (("dec_fused_d32", (32, 16)), ("dec_fused_d9", (9, 8))):
            n = 12_500_000
            z, mu = synth.latent_points(n, d, k, rank=78 + rank, device=dev)
            buf = LatentBuffer(z, n_total=n * world, group=group, exchange=exchange)
            lane = 4 * k * d + 20 * k                       # ~4Kd FMA + ~20K other lane-instructions / point

            def fused():
                buf.dec_step(mu, 1.0, 1e-3, 0)
            ms = timeit(fused, 10)
            res = {"workload": f"fused latent-buffer DEC step (assign + KL grads, centroid-only, stats all-reduced), "
                               f"{n} points per GPU x {world} GPU(s), d={d} K={k}" +
                               (" (configs[3]: 100M points over 8 GPUs)" if d == 32 else ""),
                   "points_per_s": n * world / (ms * 1e-3), "ms": ms, "algorithmic_bytes_per_point": 8 * d,
                   "hbm_gbs": 8 * d * n * world / (ms * 1e-3) / 1e9,
                   "hbm_frac": 8 * d * n / (ms * 1e-3) / 1e9 / hbm_peak,
                   "fp32_frac": 2.0 * lane * n / (ms * 1e-3) / 1e12 / fp32_peak}
            if d == 32:
                def fused_dz():
                    buf.dec_step(mu, 1.0, 1e-3, 0, want_dz=True)
                ms2 = timeit(fused_dz, 5)
                res["with_dz"] = {"ms": ms2, "points_per_s": n * world / (ms2 * 1e-3), "algorithmic_bytes_per_point": 12 * d,
                                  "hbm_frac": 12 * d * n / (ms2 * 1e-3) / 1e9 / hbm_peak}
            out[name] = res
            del z, buf
        # ---- configs[4]: one DEC_training epoch on synthetic (N,1,4,101) spectrograms, B=4096, per GPU
        nspec, bsz = 131072, 4096
        x = synth.spectrograms(nspec, rank=rank, device=dev)
        loader = synth.TensorBatches(x, bsz)
        torch.manual_seed(0)                                # identical replicas on every rank
        model = DEC(n_clusters=5).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        DEC_training(model, loader, opt, n_epochs=1, gamma=1e-3, tol=0.0, group=group)   # warm-up epoch (cuDNN autotune)
        barrier()
        t0 = time.perf_counter()
        hist = DEC_training(model, loader, opt, n_epochs=1, gamma=1e-3, tol=0.0, group=group)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        out["dec_train_epoch"] = {"workload": f"DEC_training epoch, {nspec} synthetic (1,4,101) spectrograms per GPU x {world} GPU(s), "
                                              f"B={bsz} per GPU, K=5, encoder/decoder stock torch fp32 (gradients averaged over the "
                                              "ranks), clustering path fused (configs[4])",
                                  "seconds": dt, "points_per_s": nspec * world / dt, "ms": dt * 1e3, "hbm_frac": 0.0,
                                  "loss": hist["loss"][-1] if hist["loss"] else None}
    except Exception as exc:  # pragma: no cover
        import traceback
        out["error"] = repr(exc) + " | " + traceback.format_exc()[-600:]
    return out


def gmm_cpu_baselines(torch, synth, extra):
    """Reference CPU figures for the metric's GMM half: scikit-learn's `_e_step` + `_m_step` on a
    bounded sample of the configs[2] shape from the same explicit state, and the whole `models.gmm` call of
    configs[0] (N=10k, d=9, K=8: KMeans(n_init=100) + EM) — reference formulation on the host cores beside this
    package's `models.gmm` on the GPU."""
    import warnings
    import numpy as np
    out = {}
    cores = len(os.sched_getaffinity(0))
    try:
        from sklearn.mixture import GaussianMixture as SkGM
        from sklearn.cluster import KMeans as SkKM
        n, d, k = 200_000, 9, 16
        z, _ = synth.latent_points(n, d, k, rank=77)
        X = z.numpy().astype(np.float64)                    # the reference runs float64 (models.py:66-71, 965)
        w0, mu0, cov0 = [t.numpy() for t in synth.gmm_initial_state(d, k, "cpu")]
        sk = SkGM(k, weights_init=w0, means_init=mu0, precisions_init=np.linalg.inv(cov0), max_iter=1, tol=0.0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sk.fit(X)                                       # sets the explicit state + warm-up
        best = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            _, log_resp = sk._e_step(X)
            sk._m_step(X, log_resp)
            best = min(best, time.perf_counter() - t0)
        ours = extra.get("gmm_em_iteration", {}).get("points_per_s")
        out["gmm_em_iteration_cpu"] = {"workload": f"scikit-learn _e_step + _m_step, float64, {n} points (sample of configs[2]: d=9 K=16)",
                                       "points_per_s": n / best, "seconds": best, "cores": cores, "kind": "reference (scikit-learn)",
                                       "gpu_over_cpu": (ours / (n / best)) if ours else None}
        # configs[0]: full models.gmm, reference formulation (models.py:365-413) vs this package
        n, d, k = 10_000, 9, 8
        z, _ = synth.latent_points(n, d, k, rank=5)
        Xc = z.numpy().astype(np.float64)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            km = SkKM(n_clusters=k, max_iter=1000, n_init=100, random_state=2009).fit(Xc)
            _, counts = np.unique(km.labels_, return_counts=True)
            SkGM(n_components=k, max_iter=1000, n_init=1, weights_init=counts / n, means_init=km.cluster_centers_).fit_predict(Xc)
        t_ref = time.perf_counter() - t0
        from spectrogram_cube_clustering_b200.models import gmm
        gmm(z.numpy(), k)                                   # warm-up (workspaces)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gmm(z.numpy(), k)
        torch.cuda.synchronize()
        t_gpu = time.perf_counter() - t0
        out["gmm_c1"] = {"workload": "models.gmm(z, 8): KMeans(n_init=100, max_iter=1000) seeding + full-covariance EM, "
                                     "N=10k d=9 K=8 (configs[0]); host array in, labels + centroids out",
                         "seconds": t_gpu, "reference_seconds": t_ref, "reference_cores": cores,
                         "reference": "scikit-learn KMeans + GaussianMixture exactly as Cluster/models.py:386-411 calls them",
                         "speedup": t_ref / t_gpu}
    except Exception as exc:  # pragma: no cover
        out["gmm_cpu_error"] = repr(exc)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--single-step-graphs", action="store_true", help="one graph replay per step")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--two-kernel", action="store_true",
                    help="one GPU: dec_assign + dec_target_kl_grad as two kernels instead of the one-kernel dec_step")
    ap.add_argument("--unfused", action="store_true",
                    help="three-kernel chain dec_assign -> dec_target -> dec_kl_grad(p) instead of the fused "
                         "target + KL-gradient kernel")
    ap.add_argument("--nccl", action="store_true", help="use NCCL all_reduce instead of the peer-memory exchange")
    ap.add_argument("--fused-exchange", action="store_true",
                    help="ride the exchange on the kernels (push in the producer's last CTA, pull in the consumer) "
                         "instead of the stand-alone exchange kernel")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

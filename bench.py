#!/usr/bin/env python
"""bench.py — ICP scan-pair alignments/sec on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 4096 synthetic 2-D scan pairs x 1080 points, millimetres,
3 mm range noise, 1 % invalid (0,0) returns, identity initial transform, 20 outer iterations —
per GPU (weak scaling: rank r owns pairs [r*4096, (r+1)*4096) of one global pair stream; no
collective on the data path).  One "step" = one pass of the ICP hot path
(icp_b200_estimate_batch*) over that batch.

  value : pairs/s with inputs resident in HBM (icp_b200_estimate_batch_device), CUDA events on
          the launching stream, max over ranks.
  e2e   : same metric through the reference-facing C-ABI call with HOST buffers
          (icp_b200_estimate_batch: pinned host inputs copied H2D, results copied D2H, all inside
          the timed region).
  roofline     : the fused per-pair kernel against measured HBM bandwidth (MEASURED_PEAKS.json);
                 algorithmic bytes = 16*Ns + 16*Nd + 96 per pair.
  nn_search    : the stand-alone correspondence kernel on a 100k-point map (config 3: the queries
                 15534 scans issue from their perturbed poses), queries/s and its own HBM roofline
                 (20 B/query); config3_shared_map: the ICP loop itself against that shared map.
  cpu_baseline : the oracle (C restatement of the reference, exact kd-tree) on the host cores,
                 bounded sample of the same workload.  kind "port": the reference is a Rust crate
                 and cannot be built in this image.
  config4_3d   : BASELINE config 4 — 1024 pairs (frame i, frame i+1 mod 20) of the bundled 3-D
                 recording x 16,384 points, 20 iterations — pairs/s, NN queries/s, 28 B/query roofline,
                 oracle parity sample.
  config5_1m   : (--gpus N > 1 only) BASELINE config 5 — 1,000,000 config-2 pairs in contiguous slices
                 over the N GPUs, chunks of 16,384 pairs, device-timed, max over ranks, result checksum.
  census       : inner-iteration tables and correspondence checksums of the cpu_baseline sample against
                 the oracle (0 mismatches expected: decisions are guarded, include/icp_b200.h "flags").
--impl reference times that same oracle as the reference arm (rank 0 only).
Two distinct input batches alternate between steps; each batch is 141.6 MB of points (> the
126 MB L2), so no step finds its inputs in cache.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

PAIRS_PER_GPU = 4096
BEAMS = 1080
MAX_ITER = 20
SEED = 20240807
METRIC = "ICP scan-pair alignments/sec (1080-pt 2D, 20 iterations)"
UNIT = "pairs/s"
WORKLOAD = "config2: 4096 synthetic 2D scan pairs x 1080 points per GPU, 20 outer iterations, identity init"
BYTES_PER_PAIR = 16 * BEAMS + 16 * BEAMS + 96
NN_BYTES_PER_QUERY = 20


def bench_config(world, B):
    """`config` of the JSON line: the workload only, identical in the product arm and the reference arm."""
    return {"workload": WORKLOAD if B == PAIRS_PER_GPU else f"{B} synthetic 2D scan pairs x {BEAMS} points per GPU, {MAX_ITER} outer iterations",
            "pairs_per_gpu": B, "points_per_scan": BEAMS, "outer_iterations": MAX_ITER,
            "parallelism": f"{world} x independent pair slices, no data-path collective",
            "l2": "two distinct 141.6 MB input batches alternate between steps (inputs larger than the 126 MB L2)"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel, units=None):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json),
    scaled to `units` (pairs / queries per launch) when the capture was of a different size.
    A STATIC figure from a profile, not a property of this run: bench.py says so in `traffic_source`."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            ent = json.load(f).get(kernel)
        if ent:
            b = ent.get("dram_bytes_per_launch")
            cap = ent.get("pairs") or ent.get("queries")
            if b is not None and units and cap and cap != units:
                return int(round(b * units / cap))
            return b
    return None


TRAFFIC_SOURCE = "profiles/traffic.json (static: one ncu --set full capture, scaled to this launch's units)"


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the benchmark runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, cuda_index):
        super().__init__(daemon=True)
        self.samples = []  # (t, sm_mhz, reasons bitmask)
        self.stop_flag = False
        self.ok = False
        self.sm_max = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
                except TypeError:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover - depends on the box
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                mhz = int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    rs = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.time(), mhz, rs))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self, windows):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "note": "nvml unavailable"}
        sel = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)] or self.samples
        mask = 0
        for s in sel:
            mask |= s[2]
        return {"sm_mhz": float(np.median([s[1] for s in sel])), "sm_max_mhz": self.sm_max,
                "reasons": sorted(v for k, v in self.REASONS.items() if mask & k), "samples": len(sel)}


# ----------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path (oracle port of icp_rust, exact kd-tree,
    one pair per thread on every host core) on a bounded sample of the same workload per step:
    64 pairs per core and step (1024 on 16 cores), so that a step's tail is short against its body
    and the timed region covers >= 4096 pairs."""
    if rank != 0:
        return
    import icp_rust_b200 as icp
    import oracle
    cores = host_cores()
    sample = min(PAIRS_PER_GPU, max(256, 64 * cores))
    while sample * args.steps < PAIRS_PER_GPU and sample < PAIRS_PER_GPU:
        sample *= 2
    total = sample * (args.steps + args.warmup)
    src, dst, _, _ = icp.synth_scan_pairs(total, beams=BEAMS, seed=SEED)
    off = (np.arange(sample + 1, dtype=np.uint64) * BEAMS).astype(np.uint32)
    T0 = np.tile(np.array([1., 0., 0., 1., 0., 0.]), (sample, 1))

    def step(k):
        s = src[k * sample * BEAMS:(k + 1) * sample * BEAMS]
        d = dst[k * sample * BEAMS:(k + 1) * sample * BEAMS]
        return oracle.icp_estimate_batch(2, s, off, d, off, T0, MAX_ITER, nn_mode=oracle.NN_KDTREE, nthreads=cores)

    for k in range(args.warmup):
        step(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        step(args.warmup + k)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = (f"{sample} pairs per step ({sample * args.steps} timed) of the same synthetic stream, {cores} threads, "
            "exact kd-tree; oracle port of icp_rust (Rust toolchain absent)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.gpus, PAIRS_PER_GPU),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import icp_rust_b200 as icp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: icp_rust_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = icp.Context(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()

    B, K, W = args.pairs, args.steps, args.warmup
    NBUF = 2
    host, devb = [], []
    for j in range(NBUF):
        hs = torch.empty((B * BEAMS, 2), dtype=torch.float64, pin_memory=True)
        hd = torch.empty((B * BEAMS, 2), dtype=torch.float64, pin_memory=True)
        first = ((rank + args.as_rank) * NBUF + j) * B
        _, _, off, true_T = icp.synth_scan_pairs(B, beams=BEAMS, seed=SEED, first_pair=first,
                                                 out_src=hs.numpy(), out_dst=hd.numpy())
        hoff = torch.from_numpy(off.astype(np.int32)).pin_memory()
        hT0 = torch.zeros((B, 6), dtype=torch.float64)
        hT0[:, 0] = 1.0
        hT0[:, 3] = 1.0
        hT0 = hT0.pin_memory()
        hout = torch.zeros((B, 6), dtype=torch.float64).pin_memory()
        hstat = torch.zeros(B, dtype=torch.int32).pin_memory()
        host.append(dict(src=hs, dst=hd, off=hoff, T0=hT0, out=hout, status=hstat, true_T=true_T))
        devb.append(dict(src=hs.to(dev), dst=hd.to(dev), off=hoff.to(dev), T0=hT0.to(dev),
                         out=torch.zeros((B, 6), dtype=torch.float64, device=dev),
                         status=torch.zeros(B, dtype=torch.int32, device=dev)))
    torch.cuda.synchronize()
    # a non-default stream: the C ABI treats a NULL stream as "the context's own stream", which
    # torch events recorded on the legacy default stream would not see
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    # Consecutive steps use different input and output arrays (NBUF = 2), so every step may carry
    # ICP_B200_FLAG_INDEPENDENT: its thread blocks start on the SMs the previous step's last pairs
    # leave idle.  --no-overlap times the steps strictly one after the other.
    step_flags = 0 if args.no_overlap else icp.FLAG_INDEPENDENT

    def dev_step(j, ordered=args.ordered, flags=None, c=None):
        d = devb[j % NBUF]
        (c or ctx).estimate_batch_device(dict(src=d["src"].data_ptr(), src_off=d["off"].data_ptr(), dst=d["dst"].data_ptr(),
                                       dst_off=d["off"].data_ptr(), init_T=d["T0"].data_ptr(), out_T=d["out"].data_ptr(),
                                       out_status=d["status"].data_ptr()),
                                  B, MAX_ITER, dim=2, max_src=BEAMS, max_dst=BEAMS, stream=stream.cuda_stream,
                                  synchronize=False, ordered=ordered, flags=step_flags if flags is None else flags)

    def host_step(j, c=None):
        h = host[j % NBUF]
        (c or ctx).estimate_batch_host_ptrs(dict(src=h["src"].data_ptr(), src_off=h["off"].data_ptr(), dst=h["dst"].data_ptr(),
                                          dst_off=h["off"].data_ptr(), init_T=h["T0"].data_ptr(), out_T=h["out"].data_ptr(),
                                          out_status=h["status"].data_ptr()), B, MAX_ITER, dim=2, ordered=args.ordered)

    def barrier():
        if world > 1:
            dist.barrier()

    def every_rank(x):
        if world > 1:
            t = torch.zeros(world, dtype=torch.float64, device=dev)
            t[rank] = float(x)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return [float(v) for v in t.tolist()]
        return [float(x)]

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    # ---- device-resident throughput
    for j in range(W):
        dev_step(j)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    # the K timed steps: two events around all of them (an event between two steps would serialise them)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    w0 = time.time()
    evs[0].record(stream)
    for j in range(K):
        dev_step(W + j)
    evs[K].record(stream)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    w1 = time.time()
    dev_ranks = every_rank(evs[0].elapsed_time(evs[K]) / K)
    ms_total = max_over_ranks(evs[0].elapsed_time(evs[K]))
    # the same steps strictly one after the other, an event after each: the duration of ONE launch,
    # which the roofline figures are quoted on
    for j in range(2):   # the re-run launch of a serialised step has more thread blocks: its slabs grow once
        dev_step(j, flags=0)
    torch.cuda.synchronize()
    evs[0].record(stream)
    for j in range(K):
        dev_step(W + j, flags=0)
        evs[j + 1].record(stream)
    torch.cuda.synchronize()
    each_ms = [evs[j].elapsed_time(evs[j + 1]) for j in range(K)]
    launch_ms = float(np.mean(each_ms))
    launch_ms = max_over_ranks(launch_ms)
    serial_value = world * B / (launch_ms * 1e-3)
    info = ctx.call_info()
    launches = info["kernel_launches"] * K
    for d in devb:
        assert int(d["status"].abs().sum().item()) == 0, "a scan pair was rejected on the device path"
        assert bool(torch.isfinite(d["out"]).all().item())
    value = world * B * K / (ms_total * 1e-3)
    dev_out = [d["out"].clone() for d in devb]

    # ---- the other reduction mode, briefly (reported beside the headline, not as the headline)
    other = None
    if not args.ordered:
        K2 = max(2, K // 3)
        for j in range(NBUF):
            dev_step(j, ordered=True)
        torch.cuda.synchronize()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        for j in range(K2):
            dev_step(j, ordered=True)
        o1.record(stream)
        torch.cuda.synchronize()
        oms = max_over_ranks(o0.elapsed_time(o1))
        dmax = max(float((devb[j]["out"] - dev_out[j]).abs().max().item()) for j in range(NBUF))
        other = {"flag": "ICP_B200_FLAG_ORDERED_SUMS", "value": world * B * K2 / (oms * 1e-3), "unit": UNIT,
                 "ms_per_step": oms / K2, "max_abs_transform_difference_to_default_mode": dmax}

    # ---- the same steps with EVERY outer iteration computed (ICP_B200_FIXED_POINT=0).  By default the
    # kernels write, rather than recompute, outer iterations that provably repeat (an inner solve that
    # accepted no update leaves the transform where it was: icp_fused.cuh fused_repeats); outputs are
    # bit-identical either way.  Reported beside the headline so the two can be told apart.
    os.environ["ICP_B200_FIXED_POINT"] = "0"
    try:
        ctx_full = icp.Context(local_rank)
    finally:
        os.environ.pop("ICP_B200_FIXED_POINT")
    K3 = max(2, K // 2)
    for j in range(NBUF):
        dev_step(j, c=ctx_full)
    torch.cuda.synchronize()
    full_same = all(bool(torch.equal(devb[j]["out"], dev_out[j])) for j in range(NBUF)) if not args.ordered else None
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for j in range(K3):
        dev_step(j, c=ctx_full)
    f1.record(stream)
    torch.cuda.synchronize()
    fms = max_over_ranks(f0.elapsed_time(f1))
    fixed_point = {"env": "ICP_B200_FIXED_POINT", "default": 1,
                   "every_iteration_computed": {"value": world * B * K3 / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms / K3,
                                                "bit_identical_transforms": full_same},
                   "note": "outer iterations after the first one whose inner solve accepts no update are exact repeats "
                           "(same transform, same correspondences, same solve); the default writes their outputs instead of "
                           "recomputing them.  Config 2's pairs rarely stop moving within 20 iterations; the bundled "
                           "3-D frames (config 4) stop after about six."}

    # ---- end to end through the host-pointer C-ABI call
    for j in range(NBUF):
        host_step(j)
    torch.cuda.synchronize()
    barrier()
    w2 = time.time()
    t0 = time.perf_counter()
    for j in range(K):
        host_step(j)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    w3 = time.time()
    e2e_s = max_over_ranks(e2e_s)
    hinfo = ctx.call_info()
    e2e_value = world * B * K / e2e_s
    # The same K steps from TWO host threads, each with its own context and its own pinned buffers (the
    # header's threading model: one context per thread).  The calls are the same synchronous host-pointer
    # calls; with two in flight the copies of one overlap the kernels of the other and the last pairs of
    # one call share the SMs with the first pairs of the next.
    # THREE callers: a call whose batch holds a pair the knife-edge guard re-runs returns late (the ordered
    # re-run of a lone pair shares its SM with the next batch's thread blocks and takes about a batch time), and
    # with only two calls in flight that one late call leaves the device to a single caller.  The third caller
    # reads the inputs of buffer 0 and owns its outputs.
    NCALL = 3
    ctx_b, ctx_c = icp.Context(local_rank), icp.Context(local_rank)
    callers = [ctx, ctx_b, ctx_c]
    host3 = [host[0], host[1], dict(host[0], out=torch.zeros((B, 6), dtype=torch.float64).pin_memory(),
                                    status=torch.zeros(B, dtype=torch.int32).pin_memory())]

    def caller_step(t):
        h = host3[t]
        callers[t].estimate_batch_host_ptrs(dict(src=h["src"].data_ptr(), src_off=h["off"].data_ptr(), dst=h["dst"].data_ptr(),
                                                 dst_off=h["off"].data_ptr(), init_T=h["T0"].data_ptr(), out_T=h["out"].data_ptr(),
                                                 out_status=h["status"].data_ptr()), B, MAX_ITER, dim=2, ordered=args.ordered)

    def caller(t):
        torch.cuda.set_device(dev)
        for j in range(t, K, NCALL):
            caller_step(t)

    for t in range(NCALL):
        caller_step(t)
    torch.cuda.synchronize()
    barrier()
    w4 = time.time()
    t0 = time.perf_counter()
    threads2 = [threading.Thread(target=caller, args=(t,)) for t in range(NCALL)]
    for th in threads2:
        th.start()
    for th in threads2:
        th.join()
    torch.cuda.synchronize()
    e2e2_s = time.perf_counter() - t0
    barrier()
    w5 = time.time()
    e2e2_ranks = every_rank(e2e2_s / K * 1e3)
    e2e2_s = max_over_ranks(e2e2_s)
    e2e2_value = world * B * K / e2e2_s
    same2 = bool(all(np.array_equal(host3[t]["out"].numpy(), dev_out[t % NBUF].cpu().numpy()) for t in range(NCALL)))

    # The same K steps from ONE thread through the two-half form of the call (icp_b200_estimate_batch_submit /
    # _wait), alternating the two contexts: step j+1 is enqueued before step j is waited for.
    def host_submit(j, c):
        h = host3[j % NCALL]
        c.submit_host_ptrs(dict(src=h["src"].data_ptr(), src_off=h["off"].data_ptr(), dst=h["dst"].data_ptr(),
                                dst_off=h["off"].data_ptr(), init_T=h["T0"].data_ptr(), out_T=h["out"].data_ptr(),
                                out_status=h["status"].data_ptr()), B, MAX_ITER, dim=2, ordered=args.ordered)

    def pipelined(n):
        for j in range(n):
            host_submit(j, callers[j % NCALL])
            if j >= NCALL - 1:
                callers[(j - (NCALL - 1)) % NCALL].wait()    # frees the context step j + 1 will use
        for j in range(max(0, n - (NCALL - 1)), n):
            callers[j % NCALL].wait()

    for t in range(NCALL):
        host3[t]["out"].zero_()
    pipelined(NCALL)
    torch.cuda.synchronize()
    barrier()
    w6 = time.time()
    t0 = time.perf_counter()
    pipelined(K)
    torch.cuda.synchronize()
    e2e3_s = time.perf_counter() - t0
    barrier()
    w7 = time.time()
    e2e3_ranks = every_rank(e2e3_s / K * 1e3)
    e2e3_s = max_over_ranks(e2e3_s)
    e2e3_value = world * B * K / e2e3_s
    same3 = bool(all(np.array_equal(host3[t]["out"].numpy(), dev_out[t % NBUF].cpu().numpy()) for t in range(NCALL)))
    ctx_b.close()
    ctx_c.close()
    # device-resident and host paths ran the same two batches: results must agree bit for bit
    same = bool(all(np.array_equal(host[j]["out"].numpy(), dev_out[j].cpu().numpy()) for j in range(NBUF)))

    # ---- correspondence kernel alone (config 3 shape: 100k-point map, queries near the surface)
    nn = None
    c3 = None
    c4 = None
    cpu = None
    clocks = sampler.summary([(w0, w1), (w2, w3), (w4, w5), (w6, w7)])
    sampler.stop_flag = True
    # ---- BASELINE config 5 (strong scaling) when several GPUs run: every rank takes part
    c5 = None
    if world > 1 and not args.no_config5:
        try:
            c5 = bench_config5(icp, ctx, torch, dist, dev, stream, rank, world)
        except Exception as e:  # pragma: no cover
            c5 = {"error": repr(e)}
    if rank == 0 and not args.headline_only:
        try:
            nn = bench_nn_search(icp, ctx, torch, dev, stream)
        except Exception as e:  # pragma: no cover
            nn = {"error": repr(e)}
        try:
            c3 = bench_config3(icp, ctx)
        except Exception as e:  # pragma: no cover
            c3 = {"error": repr(e)}
        try:
            c4 = bench_config4(icp, ctx, ctx_full=ctx_full)
        except Exception as e:  # pragma: no cover
            c4 = {"error": repr(e)}

    # ---- results only are gathered (48 B per pair), after the timed regions
    if world > 1:
        from icp_rust_b200.shard import gather_rows
        allT = gather_rows(dev_out[0].cpu().numpy(), world * B, device=dev)
        assert allT.shape == (world * B, 6) and np.isfinite(allT).all()

    if rank == 0:
        census = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, census = cpu_baseline(icp, ctx, host[0], B)
            cpu["single_pair_latency"] = single_pair_latency(icp, ctx)
        peak, peak_src = measured_peaks()
        achieved = BYTES_PER_PAIR * B / (launch_ms * 1e-3) / 1e9
        fp64 = fp64_roofline(icp, ctx, host[0], B, launch_ms)
        tT = host[0]["true_T"]
        oT = host[0]["out"].numpy()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "ms_per_step_every_rank": dev_ranks, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(world, B),
            "kernel_shape": {"grid": info["grid_ctas"], "block": info["block_threads"], "smem": info["smem_bytes"],
                             "launches_per_step": info["kernel_launches"],
                             "note": "one icp_lean_kernel launch + one (normally empty) ordered re-run launch of the knife-edge guard"},
            "e2e": {"value": e2e3_value, "unit": UNIT, "h2d_bytes_per_step": int(hinfo["h2d_bytes"]),
                    "d2h_bytes_per_step": int(hinfo["d2h_bytes"]), "ms_per_step": e2e3_s / K * 1e3,
                    "callers": 1, "contexts": NCALL, "bit_identical_to_device_path": same3, "ms_per_step_every_rank": e2e3_ranks,
                    "note": "K host-pointer calls (pinned host buffers in, results back in host memory) from ONE host thread "
                            "through the two-half form of the call, icp_b200_estimate_batch_submit / _wait, over three contexts: "
                            "steps j+1 and j+2 are enqueued before step j is waited for, so a step's copies and last pairs "
                            "overlap the next step's kernels, and a step that comes back late (the knife-edge guard's ordered "
                            "re-run of a flagged pair) does not leave the device short of work.  `three_callers`: K synchronous "
                            "icp_b200_estimate_batch calls from three host threads, one context each; `one_caller`: the same K "
                            "synchronous calls from a single thread, one after the other",
                    "three_callers": {"value": e2e2_value, "unit": UNIT, "ms_per_step": e2e2_s / K * 1e3,
                                      "bit_identical_to_device_path": same2, "ms_per_step_every_rank": e2e2_ranks},
                    "one_caller": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_s / K * 1e3,
                                   "bit_identical_to_device_path": same}},
            "gpu_launches": int(launches),
            "fixed_point_exit": fixed_point,
            "step_overlap": {"flag": None if args.no_overlap else "ICP_B200_FLAG_INDEPENDENT",
                             "steps_one_after_the_other": {"value": serial_value, "unit": UNIT, "ms_per_step": launch_ms,
                                                           "ms_min_max": [min(each_ms), max(each_ms)]},
                             "note": "`value` times K back-to-back asynchronous calls on one stream; with the flag a step's "
                                     "thread blocks start on the SMs the previous step's last pairs leave idle (programmatic "
                                     "dependent launch; the tenth round of 444 resident pairs of a 4096-pair step is a quarter "
                                     "full).  The roofline is quoted on one launch timed alone (launch_ms)."},
            "roofline": {"kernel": ("icp_fused_kernel" if args.ordered else "icp_lean_kernel") + f"<{info['block_threads']} threads>",
                         "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if args.ordered else profiled_traffic("icp_lean_kernel", B),
                         "traffic_source": TRAFFIC_SOURCE,
                         "peak_source": peak_src, "launch_ms": launch_ms, "algorithmic_bytes_per_launch": BYTES_PER_PAIR * B,
                         "fp64": fp64,
                         "note": "latency/FP64-chain bound by design (34.7 KB per pair, ~180 Gauss-Newton evaluations per pair); "
                                 "the FP64 pipe is the nearer ceiling (see fp64); see DESIGN.md"},
            "reduction_mode": "ordered sums (bit-exact transforms)" if args.ordered else "warp-tree reductions (default)",
            "ordered_sums_mode": other,
            "nn_search": nn,
            "config3_shared_map": c3,
            "config4_3d": c4,
            "config5_1m": c5,
            "guard": {"redo_pairs_last_host_step": int(hinfo["redo_pairs"]),
                      "note": "pairs of one 4096-pair step the knife-edge guard re-ran with ordered sums (include/icp_b200.h flags)"},
            "census": census,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "accuracy": {"median_abs_translation_error_mm": float(np.median(np.abs(oT[:, 4:] - tT[:, 4:]).max(axis=1))),
                         "median_abs_sin_error": float(np.median(np.abs(oT[:, 2] - tT[:, 2])))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time_search(torch, tgt, q, out, stream, reps):
    nq = q.shape[0]
    for _ in range(2):
        tgt.search_device(q.data_ptr(), nq, out.data_ptr(), stream=stream.cuda_stream, synchronize=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        tgt.search_device(q.data_ptr(), nq, out.data_ptr(), stream=stream.cuda_stream, synchronize=False)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_nn_search(icp, ctx, torch, dev, stream, map_points=100_000, nscan=15534, reps=5):
    """Stand-alone correspondence kernel on the config-3 map.  Main figure: the queries ICP issues —
    `nscan` scans x 1080 beams from random poses, moved by their perturbed initial transforms
    (16.8 M queries).  Second figure: map points + 20 mm Gaussian jitter in random order (no
    coherence between the lanes of a warp)."""
    import oracle
    m = icp.synth_map(map_points, seed=7)
    tgt = icp.Target(m, ctx)
    scans, off, T0 = icp.synth_map_scans(nscan, beams=BEAMS, seed=7)
    sd = torch.from_numpy(scans).to(dev).reshape(nscan, BEAMS, 2)
    Td = torch.from_numpy(T0).to(dev)
    sx_, sy_ = sd[..., 0], sd[..., 1]   # Transform::transform (transform.rs:22-24), elementwise
    q = torch.stack([Td[:, None, 0] * sx_ + Td[:, None, 1] * sy_ + Td[:, None, 4],
                     Td[:, None, 2] * sx_ + Td[:, None, 3] * sy_ + Td[:, None, 5]], dim=-1).reshape(-1, 2).contiguous()
    nq = q.shape[0]
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    md = torch.from_numpy(m).to(dev)
    pick = torch.randint(0, map_points, (nq,), device=dev, generator=g)
    qj = (md[pick] + torch.randn((nq, 2), dtype=torch.float64, device=dev, generator=g) * 20.0).contiguous()
    out = torch.empty(nq, dtype=torch.int32, device=dev)
    ms_j = _time_search(torch, tgt, qj, out, stream, reps)
    okj = bool(np.array_equal(out[:2048].cpu().numpy(), oracle.nn_search(m, qj[:2048].cpu().numpy(), oracle.NN_BRUTE)))
    ms = _time_search(torch, tgt, q, out, stream, reps)
    # spot parity on a slice (brute force is the definition)
    ok = bool(np.array_equal(out[:4096].cpu().numpy(), oracle.nn_search(m, q[:4096].cpu().numpy(), oracle.NN_BRUTE)))
    peak, peak_src = measured_peaks()
    alg = NN_BYTES_PER_QUERY * nq + 32 * map_points
    achieved = alg / (ms * 1e-3) / 1e9
    tgt.close()
    return {"metric": "NN queries/sec (2D, 100k-point map, config-3 scan queries)", "value": nq / (ms * 1e-3), "unit": "queries/s",
            "queries": nq, "map_points": map_points, "ms_per_launch": ms, "parity_sample_ok": ok and okj,
            "jittered_map_points": {"value": nq / (ms_j * 1e-3), "unit": "queries/s", "ms_per_launch": ms_j},
            "roofline": {"kernel": "nn_search_kernel<2>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": profiled_traffic("nn_search_kernel", nq), "peak_source": peak_src,
                         "l2": "268 MB of queries per launch (larger than L2)",
                         "note": "issue-bound: ~60 candidate distances per query in binary32, 2.4 of 4 issue slots, 13 of 32 lanes "
                                 "active (profiles/r01_ncu_full_summary.md); the 20 B/query stream is 2 % of HBM bandwidth"}}


def bench_config3(icp, ctx, scans_n=65536):
    """BASELINE config 3 in the ICP loop itself at full size: 65,536 scans against one shared
    100k-point map (icp_b200_target_estimate_batch, host buffers)."""
    m = icp.synth_map(100_000, seed=7)
    scans, off, T0 = icp.synth_map_scans(scans_n, beams=BEAMS, seed=7)
    tgt = icp.Target(m, ctx)
    best = None
    for _ in range(2):
        t = time.perf_counter()
        g = tgt.estimate_batch(scans, off, T0, MAX_ITER, want_inner=False, want_hash=False)
        wall = time.perf_counter() - t
        k = ctx.call_info()["kernel_ms"]
        best = (wall, k) if best is None or k < best[1] else best
    tgt.close()
    return {"workload": f"config3: {scans_n} scans x {BEAMS} points vs one shared 100000-point map, {MAX_ITER} outer iterations",
            "value": scans_n / (best[1] * 1e-3), "unit": "pairs/s", "kernel_ms": best[1],
            "host_call_pairs_per_s": scans_n / best[0], "status_ok": bool((g["status"] == 0).all()),
            "redo_pairs": int(g["redo_pairs"])}


def bench_config4(icp, ctx, pairs=1024, check=16, ctx_full=None):
    """BASELINE config 4: 1024 pairs (frame i, frame i+1 mod 20) of the bundled 3-D recording x 16,384
    points, seeded initial perturbation, 20 iterations (icp_rust_b200/workloads.py); host buffers
    through icp_b200_estimate_batch.  NN queries/s counts the 20 x 16,384 searches of every pair;
    its roofline is the correspondence stream of SURVEY.md 8(d): 28 B per 3-D query."""
    import oracle
    from icp_rust_b200.workloads import config4_pairs, CONFIG4_POINTS
    src, dst, off, T0 = config4_pairs(pairs)
    best = None
    for _ in range(2):
        t = time.perf_counter()
        g = ctx.estimate_batch(src, off, dst, off, T0, MAX_ITER, dim=3)
        wall = time.perf_counter() - t
        info = ctx.call_info()
        best = (wall, info["kernel_ms"]) if best is None or info["kernel_ms"] < best[1] else best
    n = min(pairs, check)
    o = oracle.icp_estimate_batch(3, src[:n * CONFIG4_POINTS], off[:n + 1], dst[:n * CONFIG4_POINTS], off[:n + 1], T0[:n],
                                  MAX_ITER, nn_mode=oracle.NN_KDTREE)
    peak, peak_src = measured_peaks()
    outer = computed_outer_iterations(g["inner"])
    queries = int(outer.sum()) * CONFIG4_POINTS      # searches the kernel ran (copies of a fixed point run none)
    achieved = 28.0 * queries / (best[1] * 1e-3) / 1e9
    full = None
    if ctx_full is not None:   # the same call with every outer iteration computed (ICP_B200_FIXED_POINT=0)
        gf = ctx_full.estimate_batch(src, off, dst, off, T0, MAX_ITER, dim=3)
        kf = ctx_full.call_info()["kernel_ms"]
        full = {"value": pairs / (kf * 1e-3), "unit": "pairs/s", "kernel_ms": kf,
                "nn_queries_per_s": pairs * CONFIG4_POINTS * MAX_ITER / (kf * 1e-3),
                "bit_identical_outputs": bool(np.array_equal(gf["T"], g["T"]) and np.array_equal(gf["inner"], g["inner"])
                                              and np.array_equal(gf["corr_hash"], g["corr_hash"]))}
    return {"workload": f"config4: {pairs} pairs (frame i, frame i+1 mod 20) x {CONFIG4_POINTS} 3-D points, {MAX_ITER} outer iterations",
            "value": pairs / (best[1] * 1e-3), "unit": "pairs/s", "kernel_ms": best[1], "host_call_pairs_per_s": pairs / best[0],
            "nn_queries_per_s": queries / (best[1] * 1e-3), "kernel": f"icp_stream_kernel<{info['block_threads']} threads>",
            "grid": info["grid_ctas"], "status_ok": bool((g["status"] == 0).all()), "redo_pairs": int(g["redo_pairs"]),
            "outer_iterations_computed_per_pair": float(outer.mean()), "every_iteration_computed": full,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": None,
                         "note": "28 B per 3-D query (SURVEY.md 8(d)) over the whole loop's device time: an upper-bound "
                                 "accounting, the kernel also runs the IRLS solves"},
            "oracle_parity_sample": {"pairs": n,
                                     "inner_tables_equal": bool(np.array_equal(g["inner"][:n], o["inner"])),
                                     "corr_hashes_equal": bool(np.array_equal(g["corr_hash"][:n], o["corr_hash"])),
                                     "max_abs_transform_difference": float(np.abs(g["T"][:n] - o["T"]).max())}}


def bench_config5(icp, ctx, torch, dist, dev, stream, rank, world, pairs=1_000_000, chunk=16384):
    """BASELINE config 5: 1,000,000 config-2 pairs, contiguous slices of the pair index per GPU,
    streamed through the device in chunks (generate on the host -> H2D -> timed launch), CUDA events
    on the launching stream, max over ranks; results reduced to a checksum (sum of the transforms)."""
    from icp_rust_b200.shard import pair_slice
    b0, b1 = pair_slice(pairs, rank, world)
    nthreads = max(1, host_cores() // world)
    cap = min(chunk, b1 - b0)
    d_src = torch.empty((cap * BEAMS, 2), dtype=torch.float64, device=dev)
    d_dst = torch.empty((cap * BEAMS, 2), dtype=torch.float64, device=dev)
    d_off = torch.from_numpy((np.arange(cap + 1, dtype=np.int64) * BEAMS).astype(np.int32)).to(dev)
    d_T0 = torch.zeros((cap, 6), dtype=torch.float64, device=dev)
    d_T0[:, 0] = 1.0
    d_T0[:, 3] = 1.0
    d_out = torch.empty((cap, 6), dtype=torch.float64, device=dev)
    kernel_ms = 0.0
    checksum = np.zeros(6)
    dist.barrier()
    wall0 = time.perf_counter()
    for c0 in range(b0, b1, chunk):
        n = min(chunk, b1 - c0)
        src, dst, _, _ = icp.synth_scan_pairs(n, beams=BEAMS, seed=SEED, first_pair=c0, nthreads=nthreads)
        d_src[:n * BEAMS].copy_(torch.from_numpy(src))
        d_dst[:n * BEAMS].copy_(torch.from_numpy(dst))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.estimate_batch_device(dict(src=d_src.data_ptr(), src_off=d_off.data_ptr(), dst=d_dst.data_ptr(),
                                       dst_off=d_off.data_ptr(), init_T=d_T0.data_ptr(), out_T=d_out.data_ptr()),
                                  n, MAX_ITER, dim=2, max_src=BEAMS, max_dst=BEAMS, stream=stream.cuda_stream, synchronize=False)
        e1.record(stream)
        torch.cuda.synchronize()
        kernel_ms += e0.elapsed_time(e1)
        checksum += d_out[:n].cpu().numpy().sum(axis=0)
    wall = time.perf_counter() - wall0
    t = torch.tensor([kernel_ms, wall], dtype=torch.float64, device=dev)
    cs = torch.from_numpy(checksum).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    return {"workload": f"config5: {pairs} synthetic 2D scan pairs x {BEAMS} points, {MAX_ITER} outer iterations, contiguous "
                        f"slices over {world} GPUs, chunks of {chunk} pairs",
            "scaling": "strong", "value": pairs / (float(t[0]) * 1e-3), "unit": "pairs/s",
            "device_seconds_max_over_ranks": float(t[0]) * 1e-3,
            "wall_seconds_incl_host_generation_and_copies": float(t[1]),
            "result_checksum": [float(v) for v in cs.cpu().numpy()]}


FP64_FLOPS_PER_POINT_EVAL = 80      # residual 10 + 2 rows x (weight, c, 3 jtr, 6 jtj products and sums)
FP64_FLOPS_PER_POINT_STEP = 15      # huber_error per accepted step
FP64_FLOPS_PER_POINT_OUTER = 8 + 6 * 6   # transform + ~6 exact candidate distances of the seeded search


def computed_outer_iterations(inner):
    """Outer iterations a pair actually computes with the fixed-point exit (icp_fused.cuh
    fused_repeats): up to and including the first one whose inner solve accepted no update; the
    ones after it are written as copies.  `inner`: (B, max_iter) words accepted | reason << 16."""
    idle = (np.asarray(inner) & 0xffff) == 0
    first = np.where(idle.any(axis=1), idle.argmax(axis=1) + 1, idle.shape[1])
    return first.astype(np.int64)


def fp64_roofline(icp, ctx, h, B, launch_ms, sample=256, fixed_point=True):
    """Secondary roofline of the fused kernel: algorithmic FP64 flops against the FP64
    rate measured on this pool's B200s for the reference's arithmetic (separate multiply and add, no
    FMA: tools/fp64/fp64_peak.cu -> profiles/r01_fp64_peak.json).  Evaluations per pair are counted
    from the inner-iteration table of an untimed call over the first `sample` pairs."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")) as f:
            pk = json.load(f)
        n = min(B, sample)
        off = (np.arange(n + 1, dtype=np.uint64) * BEAMS).astype(np.uint32)
        g = ctx.estimate_batch(h["src"].numpy()[:n * BEAMS], off, h["dst"].numpy()[:n * BEAMS], off, h["T0"].numpy()[:n],
                               MAX_ITER, want_inner=True, want_hash=False)
        acc = (g["inner"] & 0xffff).astype(np.int64)
        why = (g["inner"] >> 16).astype(np.int64)
        # only what the kernel computes counts: outer iterations written as copies of a fixed point
        # (one evaluation each in the reference) are left out
        outer = computed_outer_iterations(g["inner"]) if fixed_point else np.full(n, MAX_ITER)
        live = np.arange(MAX_ITER)[None, :] < outer[:, None]
        evals = float(((acc + (why != icp.BREAK_CAP)) * live).sum(axis=1).mean())
        steps = float((acc * live).sum(axis=1).mean())
        flops_pair = BEAMS * (evals * FP64_FLOPS_PER_POINT_EVAL + steps * FP64_FLOPS_PER_POINT_STEP
                              + float(outer.mean()) * FP64_FLOPS_PER_POINT_OUTER)
        tf = flops_pair * B / (launch_ms * 1e-3) / 1e12
        return {"achieved": tf, "peak": pk["fp64_dmul_dadd_tflops"], "unit": "TFLOP/s", "frac": tf / pk["fp64_dmul_dadd_tflops"],
                "peak_source": "BUILDER-measured (tools/fp64/fp64_peak.cu on this pool's B200, not in MEASURED_PEAKS.json), "
                               "separate DMUL+DADD: profiles/r01_fp64_peak.json; DFMA: %.1f" % pk["fp64_dfma_tflops"],
                "gn_evaluations_per_pair": evals, "outer_iterations_computed_per_pair": float(outer.mean()),
                "algorithmic_flops_per_pair": flops_pair,
                "note": "algorithmic flops exclude the exact selections, sqrt/div expansions and index work"}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


def single_pair_latency(icp, ctx, reps=20):
    """Latency of ONE reference call, Icp2d::new(dst).estimate(src, identity, 20) on a bundled scan pair
    (tests/golden/scans2d.npz: 001 -> 002): the oracle on one host thread against the
    single-pair entry of the C ABI (host buffers, ordered sums).  The device path is a throughput
    design; this is the honest other end of it."""
    try:
        import oracle
        z = np.load(os.path.join(ROOT, "tests", "golden", "scans2d.npz"))
        pts, off = z["points"], z["offsets"]
        src, dst = pts[off[1]:off[2]].copy(), pts[off[2]:off[3]].copy()
        I = icp.Transform.identity()
        t = time.perf_counter()
        for _ in range(reps):
            o = oracle.icp_estimate(src, dst, None, MAX_ITER, nn_mode=oracle.NN_KDTREE)
        cpu_ms = (time.perf_counter() - t) / reps * 1e3
        est = icp.Icp2d.new(dst, ctx)
        out = {}
        for ordered in (False, True):
            est.estimate(src, I, MAX_ITER, ordered=ordered)
            t = time.perf_counter()
            for _ in range(reps):
                T = est.estimate(src, I, MAX_ITER, ordered=ordered)
            out[ordered] = ((time.perf_counter() - t) / reps * 1e3, T.to_array())
        return {"pair": "bundled 2-D scans 001 -> 002 (%d / %d points), 20 iterations" % (len(src), len(dst)),
                "cpu_oracle_ms": cpu_ms, "c_abi_single_pair_ms": out[False][0],
                "max_abs_transform_difference": float(np.abs(out[False][1] - o["T"]).max()),
                "ordered_sums_ms": out[True][0], "ordered_sums_bit_identical": bool(np.array_equal(out[True][1], o["T"]))}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


def cpu_baseline(icp, ctx, h, B):
    """The oracle on every host core over the first pairs of the timed batch (~12 s), and on the same
    pairs a census: the device's inner-iteration tables and correspondence checksums against it."""
    import oracle
    cores = host_cores()
    src, dst = h["src"].numpy(), h["dst"].numpy()
    T0 = h["T0"].numpy()

    def run(n):
        off = (np.arange(n + 1, dtype=np.uint64) * BEAMS).astype(np.uint32)
        t = time.perf_counter()
        o = oracle.icp_estimate_batch(2, src[:n * BEAMS], off, dst[:n * BEAMS], off, T0[:n], MAX_ITER,
                                      nn_mode=oracle.NN_KDTREE, nthreads=cores)
        return time.perf_counter() - t, o, off

    n0 = min(B, 2 * cores)
    dt0, _, _ = run(n0)
    n = int(min(B, max(n0, 12.0 / (dt0 / n0))))
    dt, o, off = run(n)
    diff = float(np.abs(o["T"] - h["out"].numpy()[:n]).max())
    g = ctx.estimate_batch(src[:n * BEAMS], off, dst[:n * BEAMS], off, T0[:n], MAX_ITER)
    census = {"pairs": n, "inner_mismatch": int((g["inner"] != o["inner"]).any(axis=1).sum()),
              "corr_hash_mismatch": int((g["corr_hash"] != o["corr_hash"]).sum()), "redo_pairs": int(g["redo_pairs"]),
              "larger_run": "profiles/r02_census_*.json (tools/census.py, 131072 pairs)"}
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} pairs of the timed batch, {dt:.1f} s, one pair per thread, exact kd-tree",
            "gpu_max_abs_transform_difference_on_sample": diff}, census


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="scan pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ordered", action="store_true", help="ICP_B200_FLAG_ORDERED_SUMS: reference-order sums")
    ap.add_argument("--headline-only", action="store_true", help="skip the nn_search / config3 / config4 side measurements")
    ap.add_argument("--no-overlap", action="store_true",
                    help="time the device steps without ICP_B200_FLAG_INDEPENDENT (each waits for the previous one's last pair)")
    ap.add_argument("--as-rank", type=int, default=0,
                    help="diagnosis: a single-GPU run takes the input buffers rank R of a multi-GPU run would take")
    ap.add_argument("--no-config5", action="store_true", help="skip the 1M-pair strong-scaling leg of multi-GPU runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

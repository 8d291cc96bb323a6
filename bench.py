#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Powers-of-Tau -> arkworks path on B200.

Metric (BASELINE.json): G1/G2 points/sec, parse + (decompress) + on-curve + subgroup
check + arkworks re-encode.  Workload at N=1 = BASELINE configs[1]: 2^20 uncompressed
G1 tau-powers (known tau) -> canonical + on-curve + subgroup check -> ark LE.  Under
torchrun every rank owns the next contiguous 2^20-point index range of the same
tau-power section (weak scaling, no data-path collective).

One "step" = one pass of the hot path over the rank's 2^20 points:
  value : inputs already in HBM, kernel only, CUDA events on the launch stream
  e2e   : ptau_convert() on pinned HOST buffers (H2D + kernel + D2H inside the timer)
Extra legs (reported under "extra", not the headline): G1/G2 compressed (config 3),
G2 uncompressed, unchecked load, pure re-encode (HBM-bound).

`--impl reference` times the CPU restatement of the reference's algorithms
(oracle/cpu_ref.c, kind "port": the Rust reference cannot be built here) on a
bounded sample of the same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_POINTS = 20
IMAD_PER_FQMUL = 588  # 12x32-bit CIOS, lo+hi counted separately
# SURVEY.md 8d algorithmic Fq multiplications per point
FQMUL = {"g1_unc": 1030, "g2_unc": 1180, "g1_comp": 1500, "g2_comp": 2120}
METRIC = "G1 points/sec parse+on-curve+subgroup-check+ark re-encode (2^20 uncompressed tau powers per GPU)"


def sample_clocks(stop, out):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    dev = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", dev, "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 7:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.1)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm = sorted(int(float(s[0])) for s in samples)
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(samples[0][1])), "reasons": reasons,
            "samples": len(samples), "power_w_max": max(float(s[2]) for s in samples)}


def cpu_baseline(n_sample, threads, seed_tau, fast_predicates=False):
    """The oracle (C restatement of the reference's algorithms) timed on host cores.
    fast_predicates=True runs the same C code with the GPU's GLV check instead of the
    reference's multiplication by r (separates algorithmic from hardware speed-up)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref

    zu = cpu_ref.generate(1, 1, 1, seed_tau, 0, n_sample, threads)
    t0 = time.perf_counter()
    out, st = cpu_ref.convert(1, 1, zu, 3, 4 | (16 if fast_predicates else 0), threads)  # read_g1 semantics
    dt = time.perf_counter() - t0
    assert not any(st)
    return n_sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    tau = 0x1234567890ABCDEF1234567890ABCDEF
    n_sample = 1 << 14
    # size the sample to ~2-4 s per step on this host
    rate, _ = cpu_baseline(1 << 11, threads, tau)
    while n_sample / rate > 4.0 and n_sample > (1 << 11):
        n_sample >>= 1
    for _ in range(args.warmup):
        cpu_baseline(min(n_sample, 1 << 11), threads, tau)
    t_total = 0.0
    for _ in range(args.steps):
        r, dt = cpu_baseline(n_sample, threads, tau)
        t_total += dt
    value = n_sample * args.steps / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (6x64-bit Montgomery)",
        "data": "synthetic", "config": {"workload": "2^20 uncompressed G1 tau-powers (configs[1])",
                                         "sample": "%d points per step" % n_sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": "%d uncompressed G1 points per step, r-multiplication subgroup check "
                                   "(ark-ec 0.2 semantics), %d threads" % (n_sample, threads)},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_config5(args, kz, sharding, torch, dist, ctx, rank, world, dev):
    """BASELINE configs[4]: a 2^K-power synthetic setup (known tau) sharded by contiguous
    index ranges over the ranks.  Per rank: for each section of the kgz pipeline, its index
    range is generated on the GPU slab by slab (compressed zcash encoding, [s*tau^i]G) and
    pushed through the fused compressed -> strict checks -> ark kernel.  Only the convert
    kernels are timed (CUDA events); strong scaling: the total work is fixed."""
    n = 1 << args.log2_powers
    tau, alpha, beta = 0x1234567890ABCDEF1234567890ABCDEF, 0x0FEDCBA987654321, 0x13579BDF02468ACE
    secs = [("tau_g1", kz.G1, 1, 2 * n - 1), ("tau_g2", kz.G2, 1, n), ("alpha_g1", kz.G1, alpha, n),
            ("beta_g1", kz.G1, beta, n)]
    slab = 1 << 22
    d_in = torch.empty(slab * 96, dtype=torch.uint8, device=dev)
    d_out = torch.empty(slab * 192, dtype=torch.uint8, device=dev)
    status = torch.full((1,), -1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    ZC, AU = kz.FMT_ZCASH_COMPRESSED, kz.FMT_ARK_UNCOMPRESSED
    h_in = h_out = None
    if not args.no_e2e:
        h_in, h_out = kz.PinnedBuffer(slab * 96), kz.PinnedBuffer(slab * 192)
        h_in_t, h_out_t = torch.from_numpy(h_in.array), torch.from_numpy(h_out.array)
    e2e_s, h2d, d2h, e2e_launches, checked = 0.0, 0, 0, 0, set()
    # warm-up (also builds the generator tables)
    for g in (kz.G1, kz.G2):
        ctx.generate_device(g, ZC, 1, tau, 0, 4096, d_in.data_ptr(), stream=stream.cuda_stream)
        for _ in range(3):
            ctx.convert_device(g, ZC, d_in.data_ptr(), AU, d_out.data_ptr(), 4096, kz.CHECKS_STRICT, status.data_ptr(),
                               stream=stream.cuda_stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clocks), daemon=True)
    th.start()
    t_wall = time.perf_counter()
    ms = {"G1": 0.0, "G2": 0.0}
    pts = {"G1": 0, "G2": 0}
    launches = 0
    for name, g, s0, cnt in secs:
        lo, hi = sharding.shard_range(cnt, rank, world)
        for a in range(lo, hi, slab):
            c = min(slab, hi - a)
            ctx.generate_device(g, ZC, s0, tau, a, c, d_in.data_ptr(), stream=stream.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.convert_device(g, ZC, d_in.data_ptr(), AU, d_out.data_ptr(), c, kz.CHECKS_STRICT, status.data_ptr(),
                               base_index=a, stream=stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            key = "G1" if g == kz.G1 else "G2"
            ms[key] += e0.elapsed_time(e1)
            pts[key] += c
            launches += 1
            if h_in is not None:
                ri, ro = (48, 96) if g == kz.G1 else (96, 192)
                h_in_t[:c * ri].copy_(d_in[:c * ri])  # untimed: stands for the file read
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.convert(g, ZC, h_in.array[:c * ri], AU, kz.CHECKS_STRICT, out=h_out.array[:c * ro])
                e2e_s += time.perf_counter() - t0
                e2e_launches += ctx.timing()["kernel_launches"]
                h2d += c * ri
                d2h += c * ro
                if name not in checked:  # byte-compare the first slab of every section with the device-resident pass
                    checked.add(name)
                    assert torch.equal(h_out_t[:c * ro], d_out[:c * ro].cpu()), "e2e output differs (%s)" % name
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall
    stop.set()
    th.join()
    assert int(status.item()) == -1, "synthetic setup failed validation"
    if world > 1:
        dist.barrier()
    t_max = sharding.reduce_max_ms(ms["G1"] + ms["G2"])
    g1_ms, g2_ms = sharding.reduce_max_ms(ms["G1"]), sharding.reduce_max_ms(ms["G2"])
    tot_g1, tot_g2 = int(sharding.reduce_sum(pts["G1"])), int(sharding.reduce_sum(pts["G2"]))
    wall_max = sharding.reduce_max_ms(wall * 1e3)
    e2e = None
    if h_in is not None:
        e2e_ms = sharding.reduce_max_ms(e2e_s * 1e3)
        e2e = {"value": (tot_g1 + tot_g2) / (e2e_ms / 1e3), "unit": "points/s",
               "h2d_bytes_per_step": int(sharding.reduce_sum(h2d)), "d2h_bytes_per_step": int(sharding.reduce_sum(d2h)),
               "ms_per_step": e2e_ms, "note": "pinned host slabs of 2^22 points -> ptau_convert -> pinned host, max over ranks"}
        launches += e2e_launches
    if rank == 0:
        line = {
            "metric": "G1+G2 points/sec, compressed parse + sqrt decompress + subgroup check + ark re-encode "
                      "(2^%d-power setup, index-range sharded)" % args.log2_powers,
            "value": (tot_g1 + tot_g2) / (t_max / 1e3), "unit": "points/s", "n_gpus": world, "steps": 1, "warmup": 3,
            "ms_per_step": t_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (12x32-bit Montgomery, IMAD.WIDE carry chains)", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: 2^%d powers, sections tau_g1 (2N-1), tau_g2 (N), alpha_g1 (N), "
                                   "beta_g1 (N), compressed input generated on the GPU per slab" % args.log2_powers,
                       "g1_points": tot_g1, "g2_points": tot_g2, "g1_points_per_s": tot_g1 / (g1_ms / 1e3),
                       "g2_points_per_s": tot_g2 / (g2_ms / 1e3), "kernel_ms_g1_max_rank": g1_ms,
                       "kernel_ms_g2_max_rank": g2_ms, "wall_s_incl_generation": wall_max / 1e3,
                       "l2": "inputs larger than L2 (slabs of 2^22 points, 200-400 MB)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": summarize_clocks(clocks),
        }
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-points", type=int, default=LOG2_POINTS)
    ap.add_argument("--no-extra", action="store_true", help="skip the non-headline legs")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2 = headline (2^20 uncompressed G1 per GPU); config5 = 2^K-power compressed setup "
                         "sharded by index range over all ranks (strong scaling, device-resident)")
    ap.add_argument("--log2-powers", type=int, default=26, help="powers of the config5 setup")
    ap.add_argument("--no-e2e", action="store_true", help="config5: skip the host -> host leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import kzg_setup_powersoftau_b200 as kz
    from kzg_setup_powersoftau_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = kz.Context(device_ids=[local])
    if args.workload == "config5":
        run_config5(args, kz, sharding, torch, dist, ctx, rank, world, dev)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    N = 1 << args.log2_points
    tau = 0x1234567890ABCDEF1234567890ABCDEF % 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    ZU, ZC, AU, ML = kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ZCASH_COMPRESSED, kz.FMT_ARK_UNCOMPRESSED, kz.FMT_ARK_MONT_LIMBS
    STRICT = kz.CHECKS_STRICT
    first = rank * N  # this rank's index range of the tau-power section

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic input, generated on the GPU ([tau^i]G, i in this rank's range) ----
    d_in = torch.empty(N * 96, dtype=torch.uint8, device=dev)
    d_out = torch.empty(N * 96, dtype=torch.uint8, device=dev)
    ctx.generate_device(kz.G1, ZU, 1, tau, first, N, d_in.data_ptr())
    status = torch.full((1,), -1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def timed_steps(group, in_fmt, out_fmt, checks, din, dout, n, steps, warmup, all_ranks=True):
        # all_ranks=False: called by rank 0 alone (extra legs) -- no collective may be issued there
        sync = barrier if all_ranks else torch.cuda.synchronize
        for _ in range(warmup):
            ctx.convert_device(group, in_fmt, din.data_ptr(), out_fmt, dout.data_ptr(), n, checks, status.data_ptr(),
                               base_index=first, stream=stream.cuda_stream)
        sync()
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)  # evict L2 between timed iterations (outside the timed events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.convert_device(group, in_fmt, din.data_ptr(), out_fmt, dout.data_ptr(), n, checks, status.data_ptr(),
                               base_index=first, stream=stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
        sync()
        return total  # ms for `steps` launches

    clocks, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clocks), daemon=True)
    th.start()
    ms_local = timed_steps(kz.G1, ZU, AU, STRICT, d_in, d_out, N, args.steps, args.warmup)
    stop.set()
    th.join()
    assert int(status.item()) == -1, "synthetic input failed validation"
    ms = sharding.reduce_max_ms(ms_local)
    value = world * N * args.steps / (ms / 1e3)
    launches = args.steps

    # ---- e2e: host buffers through the public C-ABI call ----
    h_in = kz.PinnedBuffer(N * 96)
    h_out = kz.PinnedBuffer(N * 96)
    torch.cuda.synchronize()
    h_in_t = torch.from_numpy(h_in.array)
    h_in_t.copy_(d_in.cpu())
    for _ in range(2):
        ctx.convert(kz.G1, ZU, h_in, AU, STRICT, out=h_out)
    barrier()
    t0 = time.perf_counter()
    e2e_launches = 0
    for _ in range(args.steps):
        ctx.convert(kz.G1, ZU, h_in, AU, STRICT, out=h_out)
        e2e_launches += ctx.timing()["kernel_launches"]
    torch.cuda.synchronize()
    e2e_ms = sharding.reduce_max_ms((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e_value = world * N * args.steps / (e2e_ms / 1e3)
    assert torch.equal(torch.from_numpy(h_out.array), d_out.cpu()), "e2e output differs from device-resident output"

    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (12x32-bit Montgomery, IMAD.WIDE carry chains)", "data": "synthetic",
        "config": {"workload": "2^%d uncompressed G1 tau-powers per GPU (BASELINE configs[1]): canonical + on-curve + "
                               "GLV subgroup check + ark LE re-encode" % args.log2_points,
                   "points_per_gpu": N, "sharding": "contiguous index range per rank, no collective",
                   "l2": "flushed (256 MB fill) between timed iterations", "checks": "strict"},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": N * 96, "d2h_bytes_per_step": N * 96,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches + e2e_launches,
    }

    if rank == 0:
        line["clocks"] = summarize_clocks(clocks)
        # ---- roofline: measured IMAD peak (microbenchmark, same process) ----
        mb = {}
        for kind, name in ((0, "imad32"), (1, "imad_wide"), (2, "fq_mul")):
            t, ops = ctx.microbench(kind, 4000)
            mb[name] = ops / (t / 1e3)
        kernel_s = (ms_local / args.steps) / 1e3
        achieved = N * FQMUL["g1_unc"] * IMAD_PER_FQMUL / kernel_s
        line["roofline"] = {
            "bound": "imad", "achieved": achieved / 1e12, "peak": mb["imad32"] / 1e12, "unit": "TIMAD/s",
            "frac": achieved / mb["imad32"],
            # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel at 2^20 points, from the
            # ncu --set full capture in profiles/r01_final_ncu_full_headline_g1unc.csv (104.1 MB + 52.9 MB;
            # algorithmic 201.3 MB -- part of the output is still in L2 when the kernel ends)
            "traffic": 157.0e6 * (N / float(1 << 20)),
            "note": "integer-multiply bound, not hbm/tensor: achieved = %d Fq-mul/point x 588 IMAD (SURVEY 8d) x points / "
                    "CUDA-event launch time; peak = 32-bit IMAD microbenchmark measured in this run (148 SM x 64/clk). "
                    "HBM traffic is 192 B/point = %.2f GB/s, <0.1%% of %.0f GB/s" % (
                        FQMUL["g1_unc"], 192 * N / kernel_s / 1e9, 6552.0),
            "measured_imad_wide_per_s": mb["imad_wide"], "measured_fq_mul_per_s": mb["fq_mul"],
            # ncu sm__pipe_fmaheavy_cycles_active of this kernel (profiles/r01_final2_ncu_full_all_kernels.csv): the
            # algorithmic fraction above counts a squaring as a multiplication, the pipe counter does not
            "fma_heavy_pipe_busy_ncu": 0.857,
        }
        # ---- other kernels of the path (not the headline) ----
        extra = {}
        # the other kernels of the path, the whole-job legs and the CPU baseline are reported at N=1 only: under
        # torchrun the other ranks would sit in the closing barrier while rank 0 runs them
        if not args.no_extra and world == 1:
            def leg(name, group, in_fmt, out_fmt, checks, n, key):
                ri = kz._ffi.lib().ptau_record_size(group, in_fmt)
                ro = kz._ffi.lib().ptau_record_size(group, out_fmt)
                gen_fmt = ZU if in_fmt == AU else in_fmt
                din = torch.empty(n * ri, dtype=torch.uint8, device=dev)
                dout = torch.empty(n * ro, dtype=torch.uint8, device=dev)
                ctx.generate_device(group, gen_fmt, 1, tau, 0, n, din.data_ptr())
                if in_fmt == AU:
                    tmp = torch.empty_like(din)
                    ctx.convert_device(group, ZU, din.data_ptr(), AU, tmp.data_ptr(), n, 0, status.data_ptr())
                    torch.cuda.synchronize()
                    din = tmp
                t = timed_steps(group, in_fmt, out_fmt, checks, din, dout, n, 5, 3, all_ranks=False) / 5
                r = {"points": n, "ms": t, "points_per_s": n / (t / 1e3), "GBps": n * (ri + ro) / (t / 1e3) / 1e9}
                if key:
                    r["imad_frac"] = n * FQMUL[key] * IMAD_PER_FQMUL / (t / 1e3) / mb["imad32"]
                extra[name] = r
                del din, dout

            leg("g1_compressed_strict(config3)", kz.G1, ZC, AU, STRICT, 1 << 21, "g1_comp")
            leg("g2_compressed_strict(config3)", kz.G2, ZC, AU, STRICT, 1 << 21, "g2_comp")
            leg("g2_uncompressed_strict", kz.G2, ZU, AU, STRICT, 1 << 20, "g2_unc")
            leg("g1_decompress_only", kz.G1, ZC, ZU, 0, 1 << 21, None)
            leg("g2_decompress_only", kz.G2, ZC, ZU, 0, 1 << 21, None)
            leg("g1_load_unchecked(config4)", kz.G1, AU, ML, 0, 1 << 22, None)
            leg("g1_load_validated(config4)", kz.G1, AU, ML, STRICT, 1 << 22, None)
            leg("g1_reencode_only(hbm)", kz.G1, ZU, AU, 0, 1 << 23, None)
            assert int(status.item()) == -1
            # whole preprocess-kgz job at the ceremony's real size (2^21 powers, 604 MB response), host to host
            try:
                k = 21
                n21 = 1 << k
                L = kz._ffi.lib()
                resp = kz.PinnedBuffer(L.ptau_response_size(n21))
                setup = kz.PinnedBuffer(L.ptau_setup_size(kz.VARIANT_KGZ, n21))
                off = 64
                for g, s0, cnt in ((kz.G1, 1, 2 * n21 - 1), (kz.G2, 1, n21), (kz.G1, 7, n21), (kz.G1, 11, n21), (kz.G2, 11, 1)):
                    ln = cnt * L.ptau_record_size(g, ZC)
                    ctx.generate(g, ZC, s0, tau, 0, cnt, out=resp.array[off:off + ln])
                    off += ln
                ctx.preprocess(kz.VARIANT_KGZ, resp, n21, STRICT, out=setup)  # warm
                t0 = time.perf_counter()
                ctx.preprocess(kz.VARIANT_KGZ, resp, n21, STRICT, out=setup)
                dt = time.perf_counter() - t0
                npts = (2 * n21 - 1) + n21 + n21
                extra["preprocess_kgz_2^21_fused(host->host)"] = {
                    "points": npts, "ms": dt * 1e3, "points_per_s": npts / dt,
                    "note": "ptau_preprocess on pinned host buffers: 4.2M+2.1M compressed G1, 2.1M compressed G2, all checks"}
                # the drop-in binary on files, as a user of the reference would run it (wall clock, process start included)
                try:
                    import shutil
                    import tempfile

                    tmpd = tempfile.mkdtemp(prefix="ptau_bench_")
                    resp.array.tofile(os.path.join(tmpd, "powersoftau"))
                    exe = os.path.join(ROOT, "kzg_setup_powersoftau_b200", "bin", "preprocess-kgz")
                    import hashlib

                    hexd = hashlib.blake2b(memoryview(resp.array), digest_size=64).hexdigest()
                    runs = {}
                    for tag, flags in (("like_reference(blake2b+uncompressed_file)", ["--expect-digest", hexd]),
                                       ("no_digest_no_intermediate", ["--skip-digest", "--no-uncompressed"])):
                        for f in ("powersoftau_uncompressed", "kzg_setup"):
                            if os.path.exists(os.path.join(tmpd, f)):
                                os.remove(os.path.join(tmpd, f))
                        t0 = time.perf_counter()
                        r = subprocess.run([exe, "--dir", tmpd, "--log2-powers", str(k)] + flags,
                                           capture_output=True, text=True)
                        runs[tag] = {"wall_s": time.perf_counter() - t0, "rc": r.returncode}
                    same = open(os.path.join(tmpd, "kzg_setup"), "rb").read() == setup.array.tobytes()
                    t0 = time.perf_counter()
                    pw_l, vk_l = kz.load_kzg_setup(os.path.join(tmpd, "kzg_setup"), ctx=ctx)
                    t_load = time.perf_counter() - t0
                    t0 = time.perf_counter()
                    kz.load_kzg_setup(os.path.join(tmpd, "kzg_setup"), ctx=ctx, checks=STRICT)
                    t_loadv = time.perf_counter() - t0
                    extra["cli_preprocess_kgz_2^21(file->file)"] = {
                        "runs": runs, "output_equals_host_path": same, "load_kzg_setup_s": t_load,
                        "load_kzg_setup_validated_s": t_loadv, "tmpdir_fs": tmpd,
                        "note": "604 MB response file -> kzg_setup (604 MB), 8.4 M points, all checks; wall clock of the binary"}
                    shutil.rmtree(tmpd, ignore_errors=True)
                except Exception as e:
                    extra["cli_preprocess_kgz_2^21(file->file)"] = {"error": repr(e)}
                resp.free()
                setup.free()
            except Exception as e:  # pinned allocation of 1.2 GB may be refused on small hosts
                extra["preprocess_kgz_2^21_fused(host->host)"] = {"error": repr(e)}
            # SURVEY 8f-4 first step: KZG10 commit (MSM) over 2^20 powers, host -> host
            try:
                import numpy as np

                nc = 1 << 20
                pw = ctx.convert(kz.G1, ZU, ctx.generate(kz.G1, ZU, 1, tau, 0, nc), ML, 0)
                sc = np.random.default_rng(1).integers(0, 256, size=nc * 32, dtype=np.uint8)
                sc.reshape(nc, 32)[:, 31] &= 0x3F  # < 2^254 < r
                outp = np.zeros(104, dtype=np.uint8)
                L = kz._ffi.lib()
                for _ in range(2):
                    rcc = L.ptau_kzg_commit(ctx._h, pw.ctypes.data, sc.ctypes.data, nc, outp.ctypes.data)
                assert rcc == 0
                kms = ctx.timing()["kernel_ms"][0]
                extra["kzg10_commit_2^20(msm, kernels only)"] = {"points": nc, "ms": kms, "points_per_s": nc / (kms / 1e3),
                                                                  "launches": ctx.timing()["kernel_launches"],
                                                                  "note": "bucket method, signed 16-bit windows, counting sort by atomics, "
                                                                          "one thread per bucket"}
            except Exception as e:
                extra["kzg10_commit_2^20(msm, kernels only)"] = {"error": repr(e)}
            # KZG10::check (two Miller loops + one final exponentiation per opening, one opening per thread)
            try:
                import numpy as np

                nk = 148 * 256  # one full wave: 255 registers x 64-thread blocks = 256 resident openings per SM
                pwk = pw.reshape(-1, 104)[:32]
                g2p = ctx.convert(kz.G2, ZU, ctx.generate(kz.G2, ZU, 1, tau, 0, 2), ML, 0).reshape(2, 200)
                vk = kz.VerifierKey(g=pwk[0], gamma_g=pwk[5], h=g2p[0], beta_h=g2p[1])
                pws = kz.Powers(powers_of_g=pwk, powers_of_gamma_g=pwk[:1])
                poly = [int(x) for x in np.random.default_rng(2).integers(1, 1 << 62, size=16)]
                comm = kz.KZG10.commit(pws, poly, ctx=ctx)
                val, prf, _ = kz.KZG10.open(pws, poly, 12345, ctx=ctx)
                for _ in range(2):
                    okv = kz.KZG10.check_many(vk, np.tile(comm, (nk, 1)), [12345] * nk, [val] * nk, np.tile(prf, (nk, 1)), ctx=ctx)
                assert okv.all()
                kms = ctx.timing()["kernel_ms"][0]
                extra["kzg10_check_37888(pairings, kernel only)"] = {
                    "openings": nk, "ms": kms, "openings_per_s": nk / (kms / 1e3),
                    "note": "per opening: [v]g and [z]h scalar multiplications, two Miller loops, one final exponentiation"}
            except Exception as e:
                extra["kzg10_check_37888(pairings, kernel only)"] = {"error": repr(e)}
            hb = extra["g1_reencode_only(hbm)"]
            line["roofline_hbm"] = {"bound": "hbm", "kernel": "zcash->ark re-encode only (no checks)",
                                    "achieved": hb["GBps"], "peak": 6552.0, "unit": "GB/s",
                                    "frac": hb["GBps"] / 6552.0, "traffic": None}
        line["extra"] = extra
        # ---- CPU baseline: the reference's algorithms on this box's host cores (rank 0, N=1 only) ----
        threads = os.cpu_count() or 1
        try:
            if world > 1:
                raise RuntimeError("cpu_baseline is measured at N=1 only")
            rate1, _ = cpu_baseline(1 << 10, 1, tau)
            n_s = 1 << 13  # probe, then size the sample to ~15 s of host work (capped at the whole 2^20 workload)
            rate, dt = cpu_baseline(n_s, threads, tau)
            while n_s < N and 2 * n_s / rate <= 16.0:
                n_s *= 2
            rate, dt = cpu_baseline(n_s, threads, tau)
            line["cpu_baseline"] = {
                "value": rate, "unit": "points/s", "cores": threads, "kind": "port",
                "sample": "first %d points of the workload, C restatement of the reference's algorithms "
                          "(6x64 Montgomery, r-multiplication subgroup check), %d threads, %.1f s" % (n_s, threads, dt),
                "single_core_value": rate1,
                "same_code_with_gpu_predicates_value": cpu_baseline(min(n_s, 1 << 17), threads, tau, True)[0],
            }
        except Exception as e:  # the oracle is optional for the product, never for correctness claims
            line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": threads, "kind": "port",
                                    "sample": "not measured in this run: %s" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

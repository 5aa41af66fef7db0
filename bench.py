#!/usr/bin/env python3
"""bench.py -- coded Gbit/s decoded at fixed iterations, on N B200s, next to the host-CPU reference.

Workload (BASELINE.json configs[1] / SURVEY.md 8(d) M1): IEEE 802.3an (2048,1723) RS-LDPC
(codes/802_3/802_3_H.alist: N=2048, M=384, E=12288), normalised min-sum (the reference's
decodeNormalizedMinSum: quantizeSamples + normalizedMS), T=10 fixed iterations (the reference has no
early stop for min-sum), all-zero codeword, BPSK/AWGN at Eb/N0 = 4.0 dB, R = 0.8413.

  step        one pass of the hot path over one batch: ldpc_gpu_simulate() = Philox channel ->
              condition/quantise -> T iterations -> decisions -> error counting, all in one kernel
  value       whole-job coded Gbit/s of that step (no input to stage: samples are generated in-kernel)
  e2e         the same decoder through the reference-facing call ldpc_gpu_decode_batch() with HOST
              buffers: pinned fp32 samples H2D, kernel, packed decisions + iteration counts D2H
  roofline    algorithmic message bytes ((4E+N)*b per frame-iteration) over the kernel
              time measured with CUDA events on the launching stream, against the measured HBM peak
  cpu_baseline / --impl reference
              the reference's own object code (oracle/_ref), one process per host core

One process per GPU (torchrun for N>1); frames are sharded by frame-id range, the only collective is
the final NCCL all-reduce of the counters (ldpc_gpu_allreduce_counters).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(code="802_3_H", variant="decodeNormalizedMinSum", snr_db=4.0, R=0.8413, T=10,
                cfg=dict(Ymax=2.0, Q=6, alpha=1.25))
N_BITS, M_CHK, E_EDGES = 2048, 384, 12288


def cfg_of(abi, precision):
    return abi.default_cfg(abi.KIND_MINSUM, flags=["quantizeSamples", "normalizedMS"], num_iterations=WORKLOAD["T"],
                           precision=precision, **WORKLOAD["cfg"])


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


PROFILE_CSV = "profiles/r1p_ncu_raw_ms_rc_final.csv"      # `ncu --set full` capture of the headline kernel on this workload


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/); None if the capture is absent."""
    import csv
    path = os.path.join(ROOT, PROFILE_CSV)
    try:
        rows = list(csv.reader(open(path)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(vals[i]) * scale.get(units[i], 1.0)
        return tot
    except Exception:
        return None


def profiled_metric(name):
    """One metric of the committed ncu capture (PROFILE_CSV); None if absent."""
    import csv
    try:
        rows = list(csv.reader(open(os.path.join(ROOT, PROFILE_CSV))))
        return float(rows[2][rows[0].index(name)])
    except Exception:
        return None


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and hence its first-touch pinned host buffers) to the CPUs nearest its GPU, as listed by
    `nvidia-smi topo -m`.  torchrun leaves every rank free to float; with 8 ranks streaming host samples the
    cross-socket traffic is what the e2e number then measures.  Returns the CPU list used, or None."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        hdr = None
        for line in out.splitlines():
            cols = [c.strip() for c in line.split("\t")]
            if hdr is None and "CPU Affinity" in line:
                hdr = [c.strip() for c in line.replace("\x1b[4m", "").replace("\x1b[0m", "").split("\t")]
                continue
            if hdr and cols and cols[0].replace("\x1b[4m", "").replace("\x1b[0m", "") == "GPU%d" % local_rank:
                aff = cols[hdr.index("CPU Affinity")]
                cpus = set()
                for part in aff.split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
                cpus &= os.sched_getaffinity(0)
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    return aff
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
def cpu_reference(frames_per_proc, reps=1, cores=None):
    """The reference's CPU implementation of the path on the host cores: one process per core."""
    cores = cores or os.cpu_count() or 1
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_worker.py"), "--variant", WORKLOAD["variant"],
           "--code", WORKLOAD["code"], "--snr", str(WORKLOAD["snr_db"]), "--rate", str(WORKLOAD["R"]),
           "--iters", str(WORKLOAD["T"]), "--frames", str(frames_per_proc), "--reps", str(reps),
           "--cfg", json.dumps(WORKLOAD["cfg"])]
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd + ["--seed", str(100 + i)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for i in range(cores)]
    outs = []
    for p in procs:
        o, e = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed: " + e[-400:])
        outs.append(json.loads(o.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    fps = sum(o["frames_per_s"] for o in outs)
    return {"value": fps * N_BITS / 1e9, "unit": "Gbit/s", "cores": cores, "kind": outs[0]["kind"],
            "frames_per_s": fps, "wall_s": wall,
            "sample": "%d frames x %d rep(s) per process, %d processes, frame loop only, %s built -O2" %
                      (frames_per_proc, reps, cores, "oracle/_ref (reference object code)" if outs[0]["kind"] == "reference" else "oracle port")}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    frames = 96
    for _ in range(args.warmup):
        cpu_reference(8)
    t0 = time.perf_counter()
    res = [cpu_reference(frames) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    val = statistics.mean(r["value"] for r in res)
    line = {"impl": "reference", "metric": "coded Gbit/s decoded at fixed iters", "value": val, "unit": "Gbit/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block("cpu"),
            "cpu_baseline": {"value": val, "unit": "Gbit/s", "cores": res[0]["cores"], "kind": res[0]["kind"], "sample": res[0]["sample"]},
            "e2e": {"value": val, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def config_block(where, frames_per_step=None, e2e_frames=None):
    c = {"workload": "IEEE 802.3an (2048,1723) RS-LDPC 802_3_H.alist, normalised min-sum (decodeNormalizedMinSum: "
                     "Ymax=2.0 Q=6 alpha=1.25), T=10 fixed iterations, all-zero codeword, BPSK/AWGN Eb/N0=4.0 dB R=0.8413",
         "code": "802_3_H", "N": N_BITS, "M": M_CHK, "E": E_EDGES, "T": WORKLOAD["T"], "snr_db": WORKLOAD["snr_db"],
         "decoder": WORKLOAD["variant"]}
    if where == "gpu":
        c.update({"frames_per_step_per_gpu": frames_per_step, "e2e_frames_per_step_per_gpu": e2e_frames,
                  "parallelism": "frames sharded by frame-id range, one process per GPU, final NCCL all-reduce of counters",
                  "l2": "value: no resident input to flush (channel samples are generated in-kernel by Philox); "
                        "e2e: every step streams a fresh host batch far larger than L2"})
    return c


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per step per GPU (value)")
    ap.add_argument("--e2e-frames", type=int, default=1 << 17, help="frames per step per GPU (e2e, host buffers)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    from ldpcsimulation_b200 import abi, capi, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device visible (there is no CPU fallback)")
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference(160)

    prec = abi.PREC_F32 if args.precision == "f32" else abi.PREC_F64
    code = capi.Code(os.path.join(ROOT, "codes", "802_3", "802_3_H.alist"))
    cfg = cfg_of(abi, prec)
    dec = capi.Decoder(code, cfg, device=local)
    geo = dec.geometry()
    snr, R, F = WORKLOAD["snr_db"], WORKLOAD["R"], args.frames

    # the path's one collective: library-owned NCCL communicator, id distributed over torch.distributed
    if world > 1:
        import ctypes as C
        idbuf = (C.c_uint8 * 128)()
        if rank == 0:
            capi.check(capi.lib().ldpc_gpu_comm_unique_id(idbuf))
        t = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        idbuf = (C.c_uint8 * 128)(*t.cpu().tolist())
        capi.check(capi.lib().ldpc_gpu_comm_init(idbuf, rank, world, local))

    def step(i):
        # disjoint global frame-id ranges per (step, rank): the union over ranks is one Monte-Carlo run
        begin, n = shard.step_range(i, rank, world, F)
        return dec.simulate(snr, R, 1234, begin, n)

    launches, kernel_ms = 0, 0.0
    for i in range(args.warmup):
        r = step(i)
    if world > 1:                                       # first collective on a fresh communicator sets up its channels
        capi.check(capi.lib().ldpc_gpu_allreduce_counters(r["_cnt"], code.N, C.byref(cfg)))
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    total = None
    for i in range(args.steps):
        r = step(args.warmup + i)
        ms, nl = dec.last_timing()
        kernel_ms += ms
        launches += nl
        if total is None:
            total = dict(r.counters)
        else:
            for k in total:
                total[k] += r.counters[k]
        last = r
    if world > 1:                                       # final all-reduce of the counters, inside the timed region
        cnt = last["_cnt"]
        for k in total:
            setattr(cnt, k, total[k])
        capi.check(capi.lib().ldpc_gpu_allreduce_counters(cnt, code.N, C.byref(cfg)))
        total = cnt.as_dict()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        tw = torch.tensor([wall, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall, kernel_ms = float(tw[0]), float(tw[1])
    frames_total = F * args.steps * world
    value = frames_total * N_BITS / wall / 1e9

    # ---- e2e: reference-facing call with host buffers -------------------------------------------
    Fe = args.e2e_frames
    sigma = float(np.sqrt(10 ** (-snr / 10) / R / 2))
    gen = torch.Generator(device="cuda").manual_seed(7 + rank)
    y_dev = 1.0 + sigma * torch.randn((Fe, N_BITS), generator=gen, device="cuda", dtype=torch.float32)
    bits_host = torch.empty((Fe, N_BITS // 8), dtype=torch.uint8, pin_memory=True)
    iters_host = torch.empty((Fe,), dtype=torch.int32, pin_memory=True)
    e2e_steps = max(3, min(args.steps, 10))

    def run_e2e(torch_dtype, abi_dtype):
        y_host = torch.empty((Fe, N_BITS), dtype=torch_dtype, pin_memory=True)
        if abi_dtype == abi.DT_Q8:                                # the samples as a Q-bit converter delivers them: signed quantiser levels
            Ymax, Nq = float(WORKLOAD["cfg"]["Ymax"]), 2.0 ** WORKLOAD["cfg"]["Q"]
            a = y_dev.abs().double()
            k = torch.where(a > Ymax, torch.full_like(a, 32.0), torch.clamp(torch.floor(a * (Nq - 1.0) / (2.0 * Ymax)), min=1.0))
            y_host.copy_(torch.where(y_dev >= 0, k, -k).to(torch.int8))
        else:
            y_host.copy_(y_dev.to(torch_dtype))
        torch.cuda.synchronize()
        b = abi.Batch()
        b.n_frames, b.mem, b.y_dtype = Fe, abi.MEM_HOST, abi_dtype
        b.y, b.out_bits, b.out_iters = y_host.data_ptr(), bits_host.data_ptr(), iters_host.data_ptr()
        for _ in range(2):
            dec.decode_raw(snr, R, b)
        barrier()
        t1 = time.perf_counter()
        nl = 0
        for _ in range(e2e_steps):
            dec.decode_raw(snr, R, b)
            nl += dec.last_timing()[1]
        barrier()
        w = time.perf_counter() - t1
        if dist is not None:
            tw = torch.tensor([w], dtype=torch.float64, device="cuda")
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            w = float(tw[0])
        return Fe * e2e_steps * world * N_BITS / w / 1e9, nl, float(np.unpackbits(bits_host.numpy()).mean())

    # headline e2e: the decoder of this workload quantises its samples to Q = 6 bits (decodeNormalizedMinSum ... 2.0 6 1.25), so the
    # host hands over what a 6-bit converter delivers, one byte per sample (LDPC_GPU_DT_Q8; bit-identical to raw double samples:
    # tests/test_gpu_parity.py::test_quantiser_level_input_equals_raw_sample_input).  binary16 / fp32 samples are reported beside it.
    e2e_value, e2e_launches, ber_e2e = run_e2e(torch.int8, abi.DT_Q8)
    e2e16_value, _, ber_e2e16 = run_e2e(torch.float16, abi.DT_F16)
    e2e32_value, _, ber_e2e32 = run_e2e(torch.float32, abi.DT_F32)

    # ---- extra, labelled: the two-frames-per-thread binary16 instantiation (not the headline dtype) ----
    h2 = None
    if args.precision == "f32":
        dech = capi.Decoder(code, cfg_of(abi, abi.PREC_F16X2), device=local)
        for i in range(2):
            dech.simulate(snr, R, 1234, 10 ** 9 + i * F, F)
        barrier()
        t2 = time.perf_counter()
        hsteps = max(3, min(args.steps, 5))
        htot = None
        for i in range(hsteps):
            begin, n = shard.step_range(10 ** 6 + i, rank, world, F)
            r = dech.simulate(snr, R, 1234, begin, n)
            htot = dict(r.counters) if htot is None else {k: htot[k] + r.counters[k] for k in htot}
        barrier()
        hw = time.perf_counter() - t2
        if dist is not None:
            tw = torch.tensor([hw], dtype=torch.float64, device="cuda")
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            hw = float(tw[0])
        h2 = {"value": F * hsteps * world * N_BITS / hw / 1e9, "unit": "Gbit/s", "dtype": "f16x2",
              "fer_this_rank": htot["wordErrors"] / max(1, htot["totalWords"]), "ber_this_rank": htot["errors"] / max(1, htot["totalBits"]),
              "note": "LDPC_GPU_PREC_F16X2: two frames per thread, binary16 messages clamped to +-512; a labelled throughput "
                      "instantiation validated on decisions of converging frames and on FER confidence intervals, not the headline"}
    if world > 1:
        capi.lib().ldpc_gpu_comm_destroy()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    b_msg = 4 if prec == abi.PREC_F32 else 8
    bytes_per_frame = WORKLOAD["T"] * (4 * E_EDGES + N_BITS) * b_msg           # B_iter * T, B_io = 0 (fused channel)
    peak, peak_src = measured_peak()
    achieved = F * args.steps * bytes_per_frame / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    line = {
        "metric": "coded Gbit/s decoded at fixed iters", "value": value, "unit": "Gbit/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": config_block("gpu", F, Fe),
        "e2e": {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": Fe * N_BITS * 1,
                "d2h_bytes_per_step": Fe * (N_BITS // 8 + 4), "steps": e2e_steps,
                "api": "ldpc_gpu_decode_batch(mem=HOST, y_dtype=Q8): pinned 6-bit quantiser levels (one byte per sample) in, packed decisions + iteration counts out",
                "decoded_ber": ber_e2e,
                "fp16_samples": {"value": e2e16_value, "h2d_bytes_per_step": Fe * N_BITS * 2, "decoded_ber": ber_e2e16},
                "fp32_samples": {"value": e2e32_value, "h2d_bytes_per_step": Fe * N_BITS * 4, "decoded_ber": ber_e2e32}},
        "gpu_launches": int(launches), "e2e_gpu_launches": int(e2e_launches),
        "kernel_ms_per_step": kernel_ms / args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": profiled_traffic(),
                     "traffic_source": PROFILE_CSV + " (bytes per launch of 131072 frames)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_frame": bytes_per_frame,
                     "note": "messages never leave the SM (shared memory): DRAM traffic is nil, so algorithmic message bytes over "
                             "kernel time exceed the HBM copy peak; the binding resources are shared-memory wavefronts, the ALU pipe "
                             "and issue slots (DESIGN.md section 3, profiles/r1_summary.md)"},
        "onchip": {"profiled_l1tex_pct_of_peak": profiled_metric("l1tex__throughput.avg.pct_of_peak_sustained_active"),
                   "profiled_issue_active_pct": profiled_metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "profiled_shared_wavefronts_pct_of_peak": profiled_metric("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                   "profile": PROFILE_CSV,
                   "edge_updates_per_s": 2.0 * E_EDGES * WORKLOAD["T"] * F * args.steps / (kernel_ms * 1e-3) if kernel_ms > 0 else None,
                   "issue_slots_per_edge_iteration": (148 * 4 * 32 * 1.965e9) / (E_EDGES * WORKLOAD["T"] * F * args.steps / (kernel_ms * 1e-3)) if kernel_ms > 0 else None},
        "geometry": geo, "cpu_affinity_rank0": numa, "counters": total, "ber": total["errors"] / max(1, total["totalBits"]),
        "fer": total["wordErrors"] / max(1, total["totalWords"]),
        "clocks": clocks,
    }
    if h2 is not None:
        line["f16x2"] = h2
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

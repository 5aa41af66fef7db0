"""bench.py -- coded Gbit/s decoded at fixed iterations, on N B200s, next to the host-CPU reference.

Default workload `oms_8023` (BASELINE.json configs[1], offset variant / SURVEY.md 8(d) M1): IEEE 802.3an (2048,1723) RS-LDPC
(codes/802_3/802_3_H.alist: N=2048, M=384, E=12288), offset min-sum (the reference's decodeOffsetMinSum: quantizeSamples +
offsetMS, Ymax=1.9375 Q=5 delta=0.125 -- a dyadic lattice of step 1/16), T=10 fixed iterations (the reference has no early
stop for min-sum), all-zero codeword, BPSK/AWGN at Eb/N0 = 4.0 dB, R = 0.8413.  Decoder: LDPC_GPU_PREC_F16X2 on the exact
lattice (csrc/ldpc_ms_x2.cuh): two frames per lane in binary16, every reported decision vector certified identical to the
reference's doubles (uncertifiable frames are re-decoded by the fp64 instantiation inside the same call).

  step        one pass of the hot path over one batch: ldpc_gpu_simulate() = Philox channel -> condition/quantise ->
              T iterations -> decisions -> error counting (one kernel, plus the -- normally empty -- fp64 redo launch)
  value       whole-job coded Gbit/s of that step (no input to stage: samples are generated in-kernel)
  e2e         the same decoder through the reference-facing call ldpc_gpu_decode_batch() with HOST buffers: pinned quantiser
              levels H2D, kernel, packed decisions + iteration counts D2H; fp64 / fp32 / fp16 samples beside it
  roofline    SURVEY.md 8(d): algorithmic message bytes ((4E+N)*b per frame-iteration) over the kernel time measured with CUDA
              events on the launching stream, against the aggregate shared-memory peak 148 SM x 128 B/clk x f_SM (the state
              never leaves the SM), with the ncu-measured issue / shared-memory fractions of the committed capture
  parity_f64 / fp32 / other_workloads
              the fp64 parity instantiation and the fp32 one on the same workload; short timed runs of the other configs
  cpu_baseline* / --impl reference
              the reference's own object code (oracle/_ref): all host cores (-O2), one core (-O2), reference flags (-g, no -O)

One process per GPU (torchrun for N>1); frames are sharded by frame-id range, the only collective is the final NCCL
all-reduce of the counters (ldpc_gpu_allreduce_counters).
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

R8023 = 0.8413
WORKLOADS = {
    # name: code, reference binary, macros, cfg, T, Eb/N0, R, default precision, (Ymax, Q) of the one-byte level format or None
    "oms_8023": dict(code="802_3_H", variant="decodeOffsetMinSum", flags=["quantizeSamples", "offsetMS"],
                     cfg=dict(Ymax=1.9375, Q=5, delta=0.125), T=10, snr_db=4.0, R=R8023, precision="x2", q8=(1.9375, 5),
                     text="IEEE 802.3an (2048,1723) RS-LDPC 802_3_H.alist, offset min-sum (decodeOffsetMinSum: Ymax=1.9375 Q=5 "
                          "delta=0.125), T=10 fixed iterations, all-zero codeword, BPSK/AWGN Eb/N0=4.0 dB R=0.8413"),
    "nms_8023": dict(code="802_3_H", variant="decodeNormalizedMinSum", flags=["quantizeSamples", "normalizedMS"],
                     cfg=dict(Ymax=2.0, Q=6, alpha=1.25), T=10, snr_db=4.0, R=R8023, precision="f32", q8=(2.0, 6),
                     text="IEEE 802.3an (2048,1723) RS-LDPC 802_3_H.alist, normalised min-sum (decodeNormalizedMinSum: Ymax=2.0 Q=6 "
                          "alpha=1.25), T=10 fixed iterations, all-zero codeword, BPSK/AWGN Eb/N0=4.0 dB R=0.8413"),
    "bp_8023": dict(code="802_3_H", variant="decodeBP", flags=[], cfg=dict(), T=10, snr_db=4.0, R=R8023, precision="f32", q8=None,
                    text="802_3_H.alist, sum-product (decodeBP), T=10, Eb/N0=4.0 dB"),
    "ms_peg_t50": dict(code="PEG", variant="decodeMinSum", flags=[], cfg=dict(), T=50, snr_db=2.0, R=0.5, precision="f32", q8=None,
                       text="BASELINE configs[0]: PEGReg504x1008.alist, float min-sum (decodeMinSum), T=50, Eb/N0=2.0 dB, R=0.5"),
    "ms_dvbs2": dict(code="dvbs2", variant="decodeMinSum", flags=[], cfg=dict(), T=10, snr_db=3.0, R=0.5, precision="f32", q8=None,
                     text="BASELINE configs[3]: dvbs2_1_2.alist (N=64800, E=226799), min-sum T=10, Eb/N0=3.0 dB: messages in HBM"),
    "oms_dvbs2": dict(code="dvbs2", variant="decodeOffsetMinSum", flags=["quantizeSamples", "offsetMS"],
                      cfg=dict(Ymax=1.9375, Q=5, delta=0.125), T=10, snr_db=3.0, R=0.5, precision="x2", q8=(1.9375, 5),
                      text="BASELINE configs[3]: dvbs2_1_2.alist (N=64800, E=226799), offset min-sum (decodeOffsetMinSum: Ymax=1.9375 Q=5 delta=0.125) "
                           "T=10, Eb/N0=3.0 dB: binary16 message tiles in HBM on the exact lattice, bit-identical to the fp64 instantiation (tests/test_gpu_tileh.py)"),
    "ngdbfhw_8023": dict(code="802_3_H", variant="NGDBFhw", flags=[], cfg=dict(), T=600, snr_db=4.5, R=R8023, precision="f64", q8=None,
                         text="BASELINE configs[2]: 802_3_H.alist, NGDBFhw (integer), T<=600 with early stop, Eb/N0=4.5 dB"),
    "smngdbf_8023": dict(code="802_3_H", variant="decodeSMNGDBF", flags=None, cfg=dict(num_iterations=100, alpha=0.3, theta=-0.525, windowsize=64),
                         T=100, snr_db=4.5, R=R8023, precision="f64", q8=None,
                         text="BASELINE configs[2]: 802_3_H.alist, SM-NGDBF (decodeSMNGDBF), T<=100 with early stop, Eb/N0=4.5 dB, fp64 (parity-exact)"),
}
CODE_FILES = {"802_3_H": "802_3/802_3_H.alist", "PEG": "PEGReg504x1008/PEGReg504x1008.alist", "dvbs2": "dvbs2_1_2/dvbs2_1_2.alist"}
DTYPE_TEXT = {"x2": "f16x2", "f32": "f32", "f64": "f64"}


def cfg_of(abi, wl, precision):
    """ldpc_gpu_decoder_cfg of a workload: the reference binary's macro set + the operating point."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    prec = {"x2": abi.PREC_F16X2, "f32": abi.PREC_F32, "f64": abi.PREC_F64}[precision]
    w = WORKLOADS[wl]
    if w["flags"] is None:                                      # bit-flipping variants: macro sets live in tests/cases.py
        return cases.cfg_for(w["variant"], precision=prec, code=w["code"], **w["cfg"])
    kind = {"decodeBP": abi.KIND_BP}.get(w["variant"], abi.KIND_NGDBF_HW if w["variant"] == "NGDBFhw" else abi.KIND_MINSUM)
    kw = dict(w["cfg"])
    if kind != abi.KIND_NGDBF_HW:
        kw["num_iterations"] = w["T"]
    return abi.default_cfg(kind, flags=w["flags"], precision=prec, **kw)


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
            # nvidia-smi attaches to every GPU of the box while it starts (a few hundred ms of driver work that can stall CUDA calls of
            # all ranks): wait for its first row, so that only the periodic queries fall into the timed region
            t_end = time.time() + 3.0
            while not self.rows and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
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


PROFILE_CSV = {"x2": "profiles/r2_ncu_raw_ms_x2.csv", "f32": "profiles/r1p_ncu_raw_ms_rc_final.csv"}   # `ncu --set full` captures of the headline kernels


def profiled_traffic(csv_path):
    """dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/); None if the capture is absent."""
    import csv
    path = os.path.join(ROOT, csv_path or "-")
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


def profiled_metric(csv_path, name):
    """One metric of the committed ncu capture; None if absent."""
    import csv
    try:
        rows = list(csv.reader(open(os.path.join(ROOT, csv_path or "-"))))
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
def cpu_reference(wl, frames_per_proc, reps=1, cores=None, lib_suffix=""):
    """The reference's CPU implementation of the path on the host cores: one process per core."""
    w = WORKLOADS[wl]
    cores = cores or os.cpu_count() or 1
    over = dict(w["cfg"])
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_worker.py"), "--variant", w["variant"],
           "--code", w["code"], "--snr", str(w["snr_db"]), "--rate", str(w["R"]),
           "--iters", str(w["T"]), "--frames", str(frames_per_proc), "--reps", str(reps),
           "--cfg", json.dumps(over), "--lib-suffix", lib_suffix]
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
    nbits = outs[0]["N"]
    built = "-g, no -O (the reference's own flags, Makefile:5-6)" if lib_suffix == "_g" else "-O2"
    return {"value": fps * nbits / 1e9, "unit": "Gbit/s", "cores": cores, "kind": outs[0]["kind"],
            "frames_per_s": fps, "wall_s": wall,
            "sample": "%d frames x %d rep(s) per process, %d process(es), frame loop only, %s built %s" %
                      (frames_per_proc, reps, cores, "oracle/_ref (reference object code)" if outs[0]["kind"] == "reference" else "oracle port", built)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    frames = 96 if WORKLOADS[wl]["code"] != "dvbs2" else 4
    for _ in range(args.warmup):
        cpu_reference(wl, 8 if frames > 8 else 1)
    t0 = time.perf_counter()
    res = [cpu_reference(wl, frames) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    val = statistics.mean(r["value"] for r in res)
    line = {"impl": "reference", "metric": "coded Gbit/s decoded at fixed iters", "value": val, "unit": "Gbit/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(wl, "cpu"),
            "cpu_baseline": {"value": val, "unit": "Gbit/s", "cores": res[0]["cores"], "kind": res[0]["kind"], "sample": res[0]["sample"]},
            "e2e": {"value": val, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def config_block(wl, where, frames_per_step=None, e2e_frames=None):
    w = WORKLOADS[wl]
    c = {"workload": w["text"], "name": wl, "code": w["code"], "T": w["T"], "snr_db": w["snr_db"], "decoder": w["variant"]}
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
    ap.add_argument("--workload", default="oms_8023", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (value); 0 = per-workload default")
    ap.add_argument("--e2e-frames", type=int, default=1 << 19, help="frames per step per GPU (e2e, host buffers; the side formats use at most 2^17)")
    ap.add_argument("--precision", default=None, choices=["x2", "f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip parity_f64 / fp32 / other_workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    from ldpcsimulation_b200 import abi, capi, shard

    wl = args.workload
    W = WORKLOADS[wl]
    precision = args.precision or W["precision"]
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

    def max_over_ranks(x):
        if dist is None:
            return x
        tw = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        return float(tw[0])

    # CPU baselines first (rank 0, N=1 only), before the GPU is busy
    cpu = cpu1 = cpug = cpug1 = None
    if rank == 0 and world == 1 and not args.no_cpu:
        small = W["code"] == "dvbs2"
        cpu = cpu_reference(wl, 4 if small else 160)
        cpu1 = cpu_reference(wl, 4 if small else 320, cores=1)
        try:
            cpug = cpu_reference(wl, 2 if small else 48, lib_suffix="_g")
            cpug1 = cpu_reference(wl, 2 if small else 96, cores=1, lib_suffix="_g")
        except Exception:                                    # oracle/_ref/*_g.so not built (no reference tree at build time)
            cpug = cpug1 = None

    codes = {}

    def code_of(name):
        if name not in codes:
            codes[name] = capi.Code(os.path.join(ROOT, "codes", CODE_FILES[name]))
        return codes[name]

    code = code_of(W["code"])
    NB = code.N
    cfg = cfg_of(abi, wl, precision)
    dec = capi.Decoder(code, cfg, device=local)
    geo = dec.geometry()
    snr, R = W["snr_db"], W["R"]
    F = args.frames or {"dvbs2": 18944, "PEG": 1 << 19}.get(W["code"], 1 << 20)

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
    wall, kernel_ms = max_over_ranks(wall), max_over_ranks(kernel_ms)
    frames_total = F * args.steps * world
    value = frames_total * NB / wall / 1e9
    x2_exact, redo_frames = dec.stats()

    # ---- a short timed run of another decoder / workload (extras) ----------------------------------
    def short_value(wl2, prec2, frames, steps=3, channel=0, seed_base=10 ** 9):
        w2 = WORKLOADS[wl2]
        c2 = cfg_of(abi, wl2, prec2)
        c2.channel_mode = channel
        d2 = capi.Decoder(code_of(w2["code"]), c2, device=local)
        for i in range(2):
            d2.simulate(w2["snr_db"], w2["R"], 1234, seed_base + i * frames, frames)
        barrier()
        t2 = time.perf_counter()
        tot, kms = None, 0.0
        for i in range(steps):
            begin, n = shard.step_range(10 ** 6 + i, rank, world, frames)
            rr = d2.simulate(w2["snr_db"], w2["R"], 1234, begin, n)
            kms += d2.last_timing()[0]
            tot = dict(rr.counters) if tot is None else {k: tot[k] + rr.counters[k] for k in tot}
        barrier()
        w = max_over_ranks(time.perf_counter() - t2)
        nb = code_of(w2["code"]).N
        out = {"value": frames * steps * world * nb / w / 1e9, "unit": "Gbit/s", "dtype": DTYPE_TEXT[prec2], "frames_per_step_per_gpu": frames,
               "steps": steps, "kernel_ms_per_step": kms / steps, "fer_this_rank": tot["wordErrors"] / max(1, tot["totalWords"]),
               "ber_this_rank": tot["errors"] / max(1, tot["totalBits"]),
               "avg_iterations": tot["totalIterations"] / max(1, tot["totalWords"]), "kernel": d2.geometry()}
        return out, d2

    # ---- e2e: reference-facing call with host buffers -------------------------------------------
    # DVB-S2: two launches of 148 tiles of 64 frames (a 2048-frame batch would occupy 32 of the 148 SMs); the side formats stay at 2048 frames
    Fe = args.e2e_frames if W["code"] != "dvbs2" else (18944 if W["q8"] else 1 << 11)
    if not W["q8"]:
        Fe = min(Fe, 1 << 17)                            # no packed-level headline format: everything runs at the side formats' batch
    sigma = float(np.sqrt(10 ** (-snr / 10) / R / 2))
    gen = torch.Generator(device="cuda").manual_seed(7 + rank)
    y_dev_all = 1.0 + sigma * torch.randn((Fe, NB), generator=gen, device="cuda", dtype=torch.float32)
    bits_host = torch.empty((Fe, (NB + 7) // 8), dtype=torch.uint8, pin_memory=True)
    iters_host = torch.empty((Fe,), dtype=torch.int32, pin_memory=True)
    e2e_steps = max(3, min(args.steps, 10))
    hw_noise = None
    if W["variant"] == "NGDBFhw":                        # the parity entry takes the decoder's noise buffer from the caller
        hw_noise = torch.empty((Fe, abi.HW_QBUF), dtype=torch.float64, pin_memory=True)
        hw_noise.copy_(torch.randn((Fe, abi.HW_QBUF), generator=gen, device="cuda", dtype=torch.float64))

    Fs = min(Fe, 1 << 17 if W["code"] != "dvbs2" else 1 << 11)   # batch of the side formats (fp64 samples of 2^19 frames would pin 8.6 GB per rank)

    def run_e2e(decoder, torch_dtype, abi_dtype, steps=e2e_steps, frames=None):
        Fe = frames or Fs                                 # (shadows the headline batch size: everything below is per call)
        y_dev = y_dev_all[:Fe]
        Qb = W["q8"][1] if W["q8"] else 8
        y_host = torch.empty((Fe, NB * Qb // 8) if abi_dtype == abi.DT_QP else (Fe, NB), dtype=torch_dtype, pin_memory=True)
        if abi_dtype in (abi.DT_Q8, abi.DT_QP):                   # the samples as a Q-bit converter delivers them: signed quantiser levels
            Ymax, Nq = float(W["q8"][0]), 2.0 ** W["q8"][1]
            a = y_dev.abs().double()
            k = torch.where(a > Ymax, torch.full_like(a, 32.0), torch.clamp(torch.floor(a * (Nq - 1.0) / (2.0 * Ymax)), min=1.0))
            if abi_dtype == abi.DT_Q8:
                y_host.copy_(torch.where(y_dev >= 0, k, -k).to(torch.int8))
            else:                                                 # Q bits per sample: sign | (magnitude level - 1), saturation = all ones
                code = ((torch.clamp(k, max=2.0 ** (Qb - 1)) - 1).to(torch.uint8) | ((y_dev < 0).to(torch.uint8) << (Qb - 1)))
                sh, w8 = torch.arange(Qb, device=code.device, dtype=torch.uint8), (1 << torch.arange(8, device=code.device)).to(torch.int32)
                for f0 in range(0, Fe, 4096):                     # (bounded scratch: the bit tensor is 8 x the code tensor)
                    bits = ((code[f0:f0 + 4096].unsqueeze(-1) >> sh) & 1).reshape(-1, NB * Qb // 8, 8).to(torch.int32)
                    y_host[f0:f0 + 4096].copy_((bits * w8).sum(-1).to(torch.uint8))
        else:
            y_host.copy_(y_dev.to(torch_dtype))
        torch.cuda.synchronize()
        b = abi.Batch()
        b.n_frames, b.mem, b.y_dtype = Fe, abi.MEM_HOST, abi_dtype
        b.y, b.out_bits, b.out_iters = y_host.data_ptr(), bits_host.data_ptr(), iters_host.data_ptr()
        if hw_noise is not None:
            b.noise = hw_noise.data_ptr()
        for _ in range(2):
            decoder.decode_raw(snr, R, b)
        barrier()
        t1 = time.perf_counter()
        nl = 0
        for _ in range(steps):
            decoder.decode_raw(snr, R, b)
            nl += decoder.last_timing()[1]
        barrier()
        w = max_over_ranks(time.perf_counter() - t1)
        return Fe * steps * world * NB / w / 1e9, nl, float(np.unpackbits(bits_host.numpy()[:Fe]).mean())

    noisy_gdbf = cfg.kind == abi.KIND_GDBF and abi.noise_rows_needed(cfg) > 0
    e2e = None
    if not noisy_gdbf:                                   # (the noisy GDBF variants would stream T x N doubles of noise per frame: not a format anybody feeds)
        if W["q8"]:
            # headline e2e: the decoder quantises its samples to Q bits, so the host hands over what a Q-bit converter delivers, one byte per
            # sample (LDPC_GPU_DT_Q8; bit-identical to raw double samples: tests/test_gpu_parity.py::test_quantiser_level_input_equals_raw_sample_input)
            vp, e2e_launches, berp = run_e2e(dec, torch.uint8, abi.DT_QP, frames=Fe)
            v8, _, ber8 = run_e2e(dec, torch.int8, abi.DT_Q8, steps=3)
            e2e = {"value": vp, "unit": "Gbit/s", "h2d_bytes_per_step": Fe * NB * W["q8"][1] // 8, "d2h_bytes_per_step": Fe * ((NB + 7) // 8 + 4), "steps": e2e_steps,
                   "api": "ldpc_gpu_decode_batch(mem=HOST, y_dtype=QP): pinned %d-bit quantiser levels, bit-packed (%d bits per sample), in; packed decisions + iteration counts out" % (W["q8"][1], W["q8"][1]),
                   "decoded_ber": berp,
                   "frames_per_step_per_gpu": Fe,
                   "byte_levels": {"value": v8, "h2d_bytes_per_step": Fs * NB, "frames_per_step_per_gpu": Fs, "decoded_ber": ber8}}
        v64, l64, ber64 = run_e2e(dec, torch.float64, abi.DT_F64, steps=3)
        v32, _, ber32 = run_e2e(dec, torch.float32, abi.DT_F32, steps=3)
        v16, _, ber16 = run_e2e(dec, torch.float16, abi.DT_F16, steps=3)
        if e2e is None:
            e2e = {"value": v64, "unit": "Gbit/s", "h2d_bytes_per_step": Fs * NB * 8 + (Fs * abi.HW_QBUF * 8 if hw_noise is not None else 0),
                   "d2h_bytes_per_step": Fs * ((NB + 7) // 8 + 4), "steps": 3, "frames_per_step_per_gpu": Fs,
                   "api": "ldpc_gpu_decode_batch(mem=HOST, y_dtype=F64): pinned double samples (the reference's own sample type) in, packed decisions + iteration counts out",
                   "decoded_ber": ber64}
            e2e_launches = l64
        e2e["fp64_samples"] = {"value": v64, "h2d_bytes_per_step": Fs * NB * 8, "frames_per_step_per_gpu": Fs, "decoded_ber": ber64}
        e2e["fp32_samples"] = {"value": v32, "h2d_bytes_per_step": Fs * NB * 4, "frames_per_step_per_gpu": Fs, "decoded_ber": ber32}
        e2e["fp16_samples"] = {"value": v16, "h2d_bytes_per_step": Fs * NB * 2, "frames_per_step_per_gpu": Fs, "decoded_ber": ber16}
    else:
        e2e = {"value": None, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "api": "not measured for this workload: the parity entry of a noisy bit-flipping decoder takes T x N doubles of decoder noise per frame"}
        e2e_launches = 0

    # ---- extras on the same workload: the fp64 parity instantiation, fp32, the fast channel ----------
    extras, others = {}, {}
    if not args.no_extras and W["variant"] in ("decodeOffsetMinSum", "decodeNormalizedMinSum"):
        for name, prec2, frames in (("parity_f64", "f64", 1 << 17), ("fp32", "f32", 1 << 18), ("f16x2", "x2", 1 << 19)):
            if prec2 == precision:
                continue
            out, d2 = short_value(wl, prec2, frames)
            ve, _, _ = run_e2e(d2, torch.uint8, abi.DT_QP, steps=3)
            out["e2e"] = ve
            if prec2 == "x2":
                out["exact_lattice_kernel"], out["redo_frames"] = d2.stats()
                if not out["exact_lattice_kernel"]:
                    out["note"] = ("LDPC_GPU_PREC_F16X2 off the exact lattice: binary16 messages clamped to +-512, a labelled throughput "
                                   "instantiation (decisions of converging frames, FER confidence intervals), never the headline")
            extras[name] = out
        out, _ = short_value(wl, precision, 1 << 19, channel=abi.CHANNEL_FAST)
        extras["fast_channel"] = dict(out, note="LDPC_GPU_CHANNEL_FAST: SFU Box-Muller (lg2/sqrt/sin/cos.approx) instead of the CPU-reproducible polynomials")
    if not args.no_extras:
        for name, frames in (("nms_8023", 1 << 19), ("oms_8023", 1 << 19), ("bp_8023", 1 << 17), ("ms_peg_t50", 1 << 18), ("ms_dvbs2", 9472), ("oms_dvbs2", 18944),
                             ("ngdbfhw_8023", 1 << 19), ("smngdbf_8023", 1 << 18)):
            if name == wl:
                continue
            out, d3 = short_value(name, WORKLOADS[name]["precision"], frames)
            out["workload"] = WORKLOADS[name]["text"]
            if WORKLOADS[name]["code"] == "dvbs2":           # HBM-bound: algorithmic bytes T (4E + N) b per frame against the measured copy peak
                c3 = code_of("dvbs2")
                bm = {"x2": 2, "f32": 4, "f64": 8}[WORKLOADS[name]["precision"]]
                gbs = frames * WORKLOADS[name]["T"] * (4 * c3.E + c3.N) * bm / (out["kernel_ms_per_step"] * 1e-3) / 1e9
                out["hbm"] = {"achieved_gbs": gbs, "peak_gbs": measured_peak()[0], "frac": gbs / measured_peak()[0], "message_bytes": bm}
                if WORKLOADS[name]["precision"] == "x2":
                    out["exact_lattice_kernel"], out["redo_frames"] = d3.stats()
            others[name] = out
    if not args.no_extras:                                   # BASELINE configs[4]: non-binary GF(16) min-max, frames sharded like everything else
        nbc = capi.NbCode(os.path.join(ROOT, "codes", "NB", "gf16.reg.1536.768.alist"))
        nbd = capi.NbDecoder(nbc, 15, device=local)
        Fn = 2960
        nbd.simulate(3.5, 0.5, 1234, 0, 296)
        barrier()
        t3 = time.perf_counter()
        begin, n = shard.step_range(0, rank, world, Fn)
        rnb = nbd.simulate(3.5, 0.5, 1234, 10 ** 6 + begin, n)
        barrier()
        w3 = max_over_ranks(time.perf_counter() - t3)
        others["nb_gf16_minmax"] = {"value": Fn * world * nbc.N * nbc.m / w3 / 1e9, "unit": "Gbit/s", "dtype": "f64", "frames_per_step_per_gpu": Fn, "steps": 1,
                                    "kernel_ms_per_step": rnb.kernel_ms, "fer_this_rank": rnb.counters["wordErrors"] / max(1, rnb.counters["totalWords"]),
                                    "ber_this_rank": rnb.counters["errors"] / max(1, rnb.counters["totalBits"]),
                                    "avg_iterations": rnb.counters["totalIterations"] / max(1, rnb.counters["totalWords"]),
                                    "workload": "BASELINE configs[4]: GF(16) regular (2,4) code of 1536 symbols (codes/NB/gf16.reg.1536.768.alist, generated: the reference ships "
                                                "no GF(16) code), min-max decoding T<=15 with syndrome stop, Eb/N0=3.5 dB; parity unpinned (no runnable reference), "
                                                "checked bit for bit against the C restatement of the published algorithm"}
    if world > 1:
        capi.lib().ldpc_gpu_comm_destroy()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    # ---- roofline ------------------------------------------------------------------
    E = code.E
    b_msg = {"x2": 2, "f32": 4, "f64": 8}[precision]
    hbm_peak, peak_src = measured_peak()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9                                   # GB/s: 148 SMs x 128 B/clk x f_SM
    if W["variant"] in ("decodeMinSum", "decodeOffsetMinSum", "decodeNormalizedMinSum", "decodeBP"):
        bytes_per_frame = W["T"] * (4 * E + NB) * b_msg                           # B_iter * T, B_io = 0 (fused channel)
        iters_per_frame = W["T"]
    else:                                                                         # GDBF family: (2E)/8 bit-packed + N*b_y per executed iteration
        iters_per_frame = total["totalIterations"] / max(1, total["totalWords"])
        bytes_per_frame = iters_per_frame * (2 * E / 8 + NB * (4 if W["variant"] == "NGDBFhw" else 8))
    achieved = F * args.steps * bytes_per_frame / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    hbm_bound = W["code"] == "dvbs2"                                              # per-frame state beyond one SM: messages live in an HBM workspace
    prof = PROFILE_CSV.get(precision) if wl in ("oms_8023", "nms_8023") else None
    edge_rate = E * iters_per_frame * F * args.steps / (kernel_ms * 1e-3) if kernel_ms > 0 else None
    roof = {"bound": "hbm" if hbm_bound else "smem+issue", "achieved": achieved, "peak": hbm_peak if hbm_bound else smem_peak, "unit": "GB/s",
            "frac": (achieved / (hbm_peak if hbm_bound else smem_peak)) if achieved else None,
            "traffic": profiled_traffic(prof), "traffic_source": (prof + " (dram bytes per launch of 65536 frames)") if prof else None,
            "peak_source": peak_src if hbm_bound else "148 SMs x 128 B/clk x %.0f MHz (SM clock sampled during the timed region); SURVEY.md 8(d)" % sm_mhz,
            "algorithmic_bytes_per_frame": bytes_per_frame, "message_bytes": b_msg,
            "hbm_view": {"peak": hbm_peak, "frac": (achieved / hbm_peak) if achieved else None, "peak_source": peak_src,
                         "note": "messages never leave the SM: DRAM traffic (`traffic`) is nil, so this fraction exceeds 1 and says nothing"},
            "issue": {"lane_instr_slots_per_edge_iteration": (148 * 4 * 32 * sm_mhz * 1e6) / edge_rate if edge_rate else None,
                      "profiled_issue_active_pct": profiled_metric(prof, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "profiled_warp_instr_per_launch": profiled_metric(prof, "smsp__inst_executed.sum"),
                      "profiled_shared_wavefronts_pct_of_peak": profiled_metric(prof, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                      "profiled_l1tex_pct_of_peak": profiled_metric(prof, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
                      "profile": prof}}
    dtype_text = DTYPE_TEXT[precision]
    line = {
        "metric": "coded Gbit/s decoded at fixed iters", "value": value, "unit": "Gbit/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_text, "data": "synthetic",
        "config": config_block(wl, "gpu", F, Fe),
        "e2e": e2e,
        "gpu_launches": int(launches), "e2e_gpu_launches": int(e2e_launches),
        "kernel_ms_per_step": kernel_ms / args.steps,
        "roofline": roof,
        "geometry": geo, "cpu_affinity_rank0": numa, "counters": total, "ber": total["errors"] / max(1, total["totalBits"]),
        "fer": total["wordErrors"] / max(1, total["totalWords"]),
        "clocks": clocks,
    }
    if precision == "x2":
        line["parity"] = {"exact_lattice_kernel": bool(x2_exact), "redo_frames_fp64": int(redo_frames), "frames": int(F * (args.steps + args.warmup)),
                          "statement": ("decisions, iteration counts and counters bit-identical to the fp64 parity instantiation / the double oracle on every frame "
                                        "(tests/test_gpu_x2.py); frames the packed kernel cannot certify are re-decoded in fp64 inside the same call")
                          if x2_exact else "LDPC_GPU_PREC_F16X2 off the exact lattice: labelled approximate instantiation"}
    line.update(extras)
    if others:
        line["other_workloads"] = others
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline_1core"] = {k: cpu1[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if cpug is not None:
            line["cpu_baseline_refflags"] = {k: cpug[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline_refflags_1core"] = {k: cpug1[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py — DecompressAll throughput (uncompressed GB/s, reads/s) of the B200 path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--reads R] [--chunk C]

A "step" is one DecompressAll pass over the whole indexed gzip FASTQ (BASELINE.json
configs[1]: Generator seed 0, 10 M reads x 150 bp, gzip level 6, chunk 10,000).
  value   : uncompressed GB/s with the compressed bytes and checkpoint windows already in
            HBM (pp_job_execute only: inflate kernel -> scan -> parse kernel)
  e2e     : same metric through the C ABI with HOST buffers, every step: the kernels pull the
            compressed range and the checkpoint windows from pinned host memory over PCIe while
            they decode (PP_JOB_ZEROCOPY), and the per-chunk results are copied back to the host;
            the staged variant (cudaMemcpyAsync H2D first) is reported beside it
  roofline: the parse kernel against the measured HBM copy bandwidth (BASELINE.md §4:
            (U' + 16 R) / t_parse); the inflate kernel is branch/latency bound and is
            reported as decompressed GB/s in "inflate"
  cpu_baseline: the oracle's thread-pool DecompressAll (C restatement of the reference on the
            same zlib) on all host cores, same file, same run
Multi GPU (torchrun, one rank per GPU): chunks are independent, there is no collective on
the data path; every rank decodes its own full copy of the workload ("weak" scaling).
--impl reference times the host restatement only (the C# reference cannot run here).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CACHE = os.environ.get("PPB200_CACHE", "/tmp/ppb200_cache")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ensure_built():
    for d in ("tools", "oracle", os.path.join("parallelparsing_b200", "csrc")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, d)])


LONG = {"mean": 0.0, "sigma": 0.0, "cap": 0}  # --lognormal / --cap (BASELINE config 4: long reads)


def corpus_paths(reads, fixed, seed, chunk):
    key = f"gen_s{seed}_r{reads}_L{fixed}_gz6"
    if LONG["mean"]:
        key += f"_ln{LONG['mean']:g}_{LONG['sigma']:g}_cap{LONG['cap']}"
    d = os.path.join(CACHE, key)
    return d, os.path.join(d, "reads.fastq.gz"), os.path.join(d, f"chunk{chunk}.gzi"), os.path.join(d, "meta.json")


def make_corpus(reads, fixed, seed, chunk):
    """Generator-exact FASTQ -> one gzip member (level 6) -> IndexIO file.  Cached under /tmp."""
    import parallelparsing_b200 as pp
    d, gz_path, idx_path, meta_path = corpus_paths(reads, fixed, seed, chunk)
    os.makedirs(d, exist_ok=True)
    if not os.path.exists(gz_path):
        t = time.time()
        tmp = gz_path + f".tmp{os.getpid()}"
        gen = [os.path.join(ROOT, "tools", "_build", "ppgen"), str(reads), "--seed", str(seed)]
        if LONG["mean"]:
            gen += ["--lognormal", str(LONG["mean"]), str(LONG["sigma"])]
            if LONG["cap"]:
                gen += ["--cap", str(LONG["cap"])]
        elif fixed:
            gen += ["--fixed", str(fixed)]
        p1 = subprocess.Popen(gen, stdout=subprocess.PIPE)
        p2 = subprocess.Popen([os.path.join(ROOT, "tools", "_build", "ppgzip"), "-l", "6", "-", tmp], stdin=p1.stdout)
        p1.stdout.close()
        if p2.wait() != 0 or p1.wait() != 0:
            raise RuntimeError("corpus generation failed")
        os.replace(tmp, gz_path)
        log(f"[bench] corpus {reads} reads -> {os.path.getsize(gz_path)/1e6:.1f} MB gz in {time.time()-t:.1f}s")
    if not os.path.exists(idx_path):
        t = time.time()
        # records longer than 32 768 B make the reference's CreateIndex throw (quirk H2): lifted only for
        # uncapped long reads, a documented extension
        ix = pp.Core.BuildDeflateIndex(gz_path, chunk, lift_record_cap=bool(LONG["mean"]) and not LONG["cap"])
        tmp = idx_path + f".tmp{os.getpid()}"
        pp.IndexIO.Serialize(ix, tmp)
        os.replace(tmp, idx_path)
        log(f"[bench] CreateIndex chunk={chunk}: {ix.Count} points in {time.time()-t:.1f}s")
    return gz_path, idx_path


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for i, nm in enumerate(names):
                    if s[3 + i].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_ncu_traffic(reads, chunk):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
    captures (profiles/ncu_traffic.json), for the workload they were taken on; else nothing."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        for e in json.load(open(p)):
            if e.get("reads") == reads and e.get("chunk") == chunk:
                return e.get("dram_bytes_per_launch", {})
    except Exception:
        pass
    return {}


def bind_to_gpu_numa_node(gpu):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers
    it allocates and fills (first touch) are local to that GPU's PCIe root.  Best effort: any failure
    leaves the affinity alone."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(gpu).pci_bus_id
        dom = torch.cuda.get_device_properties(gpu).pci_domain_id
        dev = torch.cuda.get_device_properties(gpu).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            log(f"[bench] rank on GPU {gpu}: NUMA node {node}, {len(cpus)} CPUs")
    except Exception as e:  # noqa: BLE001
        log(f"[bench] NUMA binding skipped: {e}")


def cpu_reference(gz, idx_path, threads, steps, warmup):
    """The host restatement of the reference's parallel DecompressAll, all cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    ox = O.OracleIndex.deserialize(idx_path)
    times, recs, nbytes = [], 0, 0
    for i in range(warmup + steps):
        t = time.perf_counter()
        recs, nbytes = O.decompress_all_mt(gz, ox, threads=threads)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    if recs < 0:
        raise RuntimeError(f"oracle failed rc={recs}")
    return float(np.mean(times)), recs, nbytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--chunk", type=int, default=10_000)
    ap.add_argument("--fixed-len", type=int, default=150)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lognormal", type=float, nargs=2, metavar=("MEAN", "SIGMA"),
                    help="read lengths lognormal with this mean (BASELINE config 4: --lognormal 10000 0.5 --chunk 1000)")
    ap.add_argument("--cap", type=int, default=0, help="clamp read lengths (16000 keeps records reference-legal, H2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.lognormal:
        LONG.update(mean=args.lognormal[0], sigma=args.lognormal[1], cap=args.cap)
        args.fixed_len = 0
    lens = (f"lognormal lengths (mean {LONG['mean']:g} bp, sigma {LONG['sigma']:g}"
            f"{', capped at ' + str(LONG['cap']) if LONG['cap'] else ', uncapped: record cap lifted'})"
            if LONG["mean"] else f"{args.fixed_len}bp")
    workload = (f"Generator seed {args.seed}, {args.reads} reads x {lens} single-end, gzip -6 "
                f"(one member), chunk {args.chunk}")
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        ensure_built()
        gz_path, idx_path = make_corpus(args.reads, args.fixed_len, args.seed, args.chunk)
        gz = np.fromfile(gz_path, np.uint8)
        dt, recs, nbytes = cpu_reference(gz, idx_path, cores, max(args.steps, 1), min(args.warmup, 1))
        val = nbytes / dt / 1e9
        line = {
            "impl": "reference", "metric": "DecompressAll uncompressed GB/s", "value": val, "unit": "GB/s",
            "reads_per_s": recs / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "records": recs, "uncompressed_bytes": nbytes,
                       "note": "C restatement of the reference's thread-pool DecompressAll on the same system zlib; "
                               "the C# reference cannot run in this image (no dotnet)"},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": "whole workload, every step"},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm
    # stdout carries exactly ONE JSON line: anything a library prints while we run (NCCL's version
    # banner, build output) is sent to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if local_rank == 0:
        ensure_built()
        make_corpus(args.reads, args.fixed_len, args.seed, args.chunk)
    barrier()
    if world > 1:
        bind_to_gpu_numa_node(local_rank)  # several GPUs pull from host memory at once: keep each rank's buffers local
    import parallelparsing_b200 as pp
    from parallelparsing_b200 import _lib
    L = pp.lib()
    gz_path, idx_path = corpus_paths(args.reads, args.fixed_len, args.seed, args.chunk)[1:3]
    gz_np = np.fromfile(gz_path, np.uint8)
    gz, gz_ptr = pp.pinned_copy(gz_np)           # pinned host memory: the e2e source buffer
    ix = pp.IndexIO.Deserialize(idx_path)
    dev = pp.Device(local_rank)
    job = pp.Job(dev, ix, gz.size, 0, -1, zero_copy=False)     # staged: H2D copies, kernels read HBM
    job_zc = pp.Job(dev, ix, gz.size, 0, -1, zero_copy=True)   # pull: kernels read pinned host memory

    def sync():
        torch.cuda.synchronize()

    # correctness gate before timing: one full pass per mode; totals must agree
    info = job.run(gz)
    if info.status != 0:
        raise SystemExit(f"bench.py: DecompressAll failed status={info.status}")
    U, R, Us = info.total_bytes, info.total_records, info.scanned_bytes
    n_chunks = info.n_chunks
    h2d_bytes = info.h2d_bytes
    iz = job_zc.run(gz)
    if iz.status != 0 or (iz.total_bytes, iz.total_records) != (U, R):
        raise SystemExit("bench.py: zero-copy DecompressAll disagrees with the staged run")
    # size-independent properties at the full workload size: every byte of the stream is produced
    # (the end sentinel's Output), every chunk is exactly to.Output-from.Output long, no read is lost
    outs = ix.scalars()[0]
    if U != int(outs[-1] - outs[0]) or R < args.reads:
        raise SystemExit(f"bench.py: DecompressAll produced {U} bytes / {R} records, expected {int(outs[-1] - outs[0])} / >= {args.reads}")
    for k in (0, n_chunks // 2, n_chunks - 1):
        if job.chunk(k).inflated != int(outs[k + 1] - outs[k]):
            raise SystemExit(f"bench.py: chunk {k} has the wrong length")
    # checksum of checksums: the gzip trailer holds CRC-32 and length of the whole uncompressed stream;
    # the concatenation of all inflated chunks (pull-mode run) must reproduce both
    import zlib
    allb = job_zc.all_bytes()
    crc = zlib.crc32(memoryview(allb)) & 0xffffffff
    want_crc, want_len = int(gz_np[-8:-4].view("<u4")[0]), int(gz_np[-4:].view("<u4")[0])
    if crc != want_crc or (allb.size & 0xffffffff) != want_len:
        raise SystemExit(f"bench.py: CRC-32/ISIZE of the inflated stream {crc:#x}/{allb.size} != gzip trailer {want_crc:#x}/{want_len}")
    del allb

    sampler = ClockSampler(local_rank)
    sampler.start()

    def timed(step, steps, warmup):
        """warmup, barrier+sync, K steps, sync; CUDA events bracket the K steps on the default stream
        after the library's own stream has been joined (sync() on both sides)."""
        for _ in range(warmup):
            step()
        sync()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        sync()
        return (time.perf_counter() - t0) / steps

    # --- kernel-only: inputs resident in HBM (pp_job_execute: inflate -> scan -> parse) -------------
    job.upload(gz_ptr)
    sync()
    t_dev = timed(job.execute, args.steps, args.warmup)
    job.download()
    info = job.info()
    t_inflate, t_parse, t_scan = info.inflate_ms * 1e-3, info.parse_ms * 1e-3, info.scan_ms * 1e-3
    launches_per_step = info.launches

    # --- end to end through the C ABI, host buffers every step ------------------------------------
    # (a) pull mode: the kernels read the compressed range and the windows from pinned host memory
    #     (TMA bulk copies over PCIe) while they decode; per-chunk results come back to the host.
    def step_zc():
        job_zc.upload(gz_ptr)
        job_zc.execute()
        job_zc.download()     # synchronises: per-chunk status, counts, record bases on the host
    t_e2e = timed(step_zc, args.steps, 2)
    d2h_bytes = job_zc.info().d2h_bytes

    # (b) staged mode: cudaMemcpyAsync of the same bytes, then the kernels, then the results
    def step_staged():
        job.upload(gz_ptr)
        job.execute()
        job.download()
    t_e2e_staged = timed(step_staged, max(2, args.steps // 2), 1)
    barrier()
    clocks = sampler.finish()

    # --- optional: also bring every record's line offsets to the host ---------------------------
    t0 = time.perf_counter()
    step_zc()
    ls = job_zc.line_starts()
    t_e2e_offsets = time.perf_counter() - t0
    del ls
    info2 = job.info()

    # max over ranks
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, t_inflate, t_parse, t_e2e_offsets, t_e2e_staged], device="cuda",
                          dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_inflate, t_parse, t_e2e_offsets, t_e2e_staged = [float(x) for x in tt.tolist()]

    if rank == 0:
        peak, peak_src = load_peaks()
        b_parse = Us + 16 * R                      # bytes scanned + four u32 line starts / record
        b_inflate = info2.compressed_bytes + 32768 * n_chunks + U
        traffic = load_ncu_traffic(args.reads, args.chunk)
        line = {
            "metric": "DecompressAll uncompressed GB/s", "value": world * U / t_dev / 1e9, "unit": "GB/s",
            "reads_per_s": world * R / t_dev,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "chunks": n_chunks, "records": R, "uncompressed_bytes": U,
                       "compressed_bytes": info2.compressed_bytes, "per_gpu": "full workload per rank",
                       "l2": f"inputs larger than L2 ({info2.compressed_bytes / 1e9:.1f} GB compressed in, "
                             f"{U / 1e9:.1f} GB inflated out per step vs 126 MB L2)"},
            "e2e": {"value": world * U / t_e2e / 1e9, "unit": "GB/s", "reads_per_s": world * R / t_e2e,
                    "ms_per_step": t_e2e * 1e3, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "mode": "pull: kernels read the compressed range + checkpoint windows from pinned host memory",
                    "staged_copy": {"value": world * U / t_e2e_staged / 1e9, "unit": "GB/s",
                                    "ms_per_step": t_e2e_staged * 1e3,
                                    "mode": "cudaMemcpyAsync H2D, then kernels, then D2H of results"},
                    "with_line_offsets_to_host": {"value": world * U / t_e2e_offsets / 1e9, "unit": "GB/s",
                                                  "d2h_bytes_per_step": d2h_bytes + 16 * R}},
            "roofline": {"kernel": "pp_parse_kernel", "bound": "hbm", "achieved": b_parse / t_parse / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": b_parse / t_parse / 1e9 / peak,
                         "traffic": traffic.get("pp_parse_kernel"),
                         "algorithmic_bytes": b_parse, "ms": t_parse * 1e3, "peak_source": peak_src,
                         "timing": "CUDA events on the library stream around the kernel, last timed step",
                         "peak_kind": "device COPY bandwidth (read + write); this kernel reads ~25x more than it "
                                      "writes, and a read-only stream can run above that figure (long reads: "
                                      "frac > 1)"},
            "inflate": {"kernel": "pp_inflate_kernel",
                        "bound": "instruction issue / shared-memory latency (Huffman decode + LZ77 resolve), not HBM",
                        "decompressed_gbs": U / t_inflate / 1e9, "ms": t_inflate * 1e3,
                        "algorithmic_bytes": b_inflate, "hbm_frac": b_inflate / t_inflate / 1e9 / peak,
                        "traffic": traffic.get("pp_inflate_kernel"),
                        "share_of_step": t_inflate / t_dev},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            dt, recs, nbytes = cpu_reference(gz_np, idx_path, cores, 2, 1)
            assert recs == R and nbytes == U, (recs, R, nbytes, U)
            line["cpu_baseline"] = {"value": nbytes / dt / 1e9, "unit": "GB/s", "reads_per_s": recs / dt,
                                    "cores": cores, "kind": "port",
                                    "sample": "whole workload, 2 timed passes after 1 warm-up (oracle thread pool)"}
            # the reference's serial baseline (SimpleDecompressor: GZipStream + line parser), one thread,
            # on a bounded prefix of the same file (a truncated stream ends in an error: only the bytes
            # it got through are used)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            sample = gz_np[: min(gz_np.size, 192 << 20)]
            t0 = time.perf_counter()
            _, nb = O.naive_count(sample)
            dtn = time.perf_counter() - t0
            line["cpu_baseline"]["naive_serial"] = {
                "value": nb / dtn / 1e9, "unit": "GB/s", "reads_per_s": nb / dtn / (U / R), "cores": 1, "kind": "port",
                "sample": f"first {sample.size >> 20} MiB of the compressed file ({nb} bytes inflated and parsed)"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    job_zc.free()
    job.free()
    L.pp_host_free(gz_ptr)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

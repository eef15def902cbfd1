#!/usr/bin/env python
"""bench.py — DecompressAll throughput (uncompressed GB/s, reads/s) of the B200 path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--reads R] [--chunk C]

A "step" is one DecompressAll pass over ONE indexed gzip FASTQ.  At N = 1 that file is BASELINE.json
configs[1] (Generator seed 0, 10 M reads x 150 bp, gzip level 6, chunk 10,000).  At N GPUs the file
holds N x R reads and is PARTITIONED: the chunk list is cut into N contiguous ranges of near-equal
compressed size (pp_partition_chunks), one per rank/GPU, and every rank holds and moves only ITS
compressed byte range and ITS checkpoint windows (weak scaling: the work per GPU is fixed).  There is
no collective on the data path (chunks are independent); NCCL carries only barriers and the max over
ranks of the timings.
  value   : uncompressed GB/s with the compressed bytes and checkpoint windows already in HBM
            (pp_job_execute only: inflate kernel -> scan -> parse kernel)
  e2e     : same metric through the C ABI with HOST buffers, every step: compressed range + windows from
            pinned host memory to the GPU, kernels, per-chunk results back to the host.  Two ways are
            timed and the faster is the headline: "pipelined" (copy-engine H2D in pieces, overlapped with
            the kernels, PP_JOB_PIPELINE) and "pull" (the kernels read pinned host memory themselves,
            PP_JOB_ZEROCOPY); both ship the windows compressed (PP_JOB_COMPACT_WINDOWS).  Beside it:
            one_call (cold pp_decompress_all + pp_job_free per step), with_line_offsets_to_host and
            with_bytes_to_host (the whole inflated stream streamed to pinned host memory while decoding).
  roofline: the parse kernel against the measured HBM copy bandwidth ((U' + 16 R) / t_parse); the
            inflate kernel is instruction-issue bound and is reported as decompressed GB/s in "inflate"
  cpu_baseline: the oracle's thread-pool DecompressAll (C restatement of the reference on the same
            zlib) on all host cores, same file, same run (N = 1 only)
--impl reference times the host restatement only (the C# reference cannot run here); it touches
nothing of the product: corpus by tools/, index by the oracle's own CreateIndex.
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
SHARD_READS = 10_000_000   # files above this are made of shards generated in parallel, seeds seed, seed+1, ...


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_native(dirs):
    for d in dirs:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, d)])


LONG = {"mean": 0.0, "sigma": 0.0, "cap": 0}  # --lognormal / --cap (BASELINE config 4: long reads)


def corpus_paths(reads, fixed, seed, chunk):
    key = f"gen_s{seed}_r{reads}_L{fixed}_gz6"
    if LONG["mean"]:
        key += f"_ln{LONG['mean']:g}_{LONG['sigma']:g}_cap{LONG['cap']}"
    d = os.path.join(CACHE, key)
    return d, os.path.join(d, "reads.fastq.gz"), os.path.join(d, f"chunk{chunk}.gzi")


def gen_cmd(reads, seed, fixed):
    cmd = [os.path.join(ROOT, "tools", "_build", "ppgen"), str(reads), "--seed", str(seed)]
    if LONG["mean"]:
        cmd += ["--lognormal", str(LONG["mean"]), str(LONG["sigma"])]
        if LONG["cap"]:
            cmd += ["--cap", str(LONG["cap"])]
    elif fixed:
        cmd += ["--fixed", str(fixed)]
    return cmd


def make_gz(reads, fixed, seed, chunk):
    """Generator-exact FASTQ -> ONE gzip member (level 6).  Cached under /tmp.  Files of more than
    SHARD_READS reads are the concatenation of shards of SHARD_READS reads with seeds seed, seed+1, ...
    (the .NET PRNG is serial; shards are generated in parallel) — still one stream, one member."""
    d, gz_path, _ = corpus_paths(reads, fixed, seed, chunk)
    os.makedirs(d, exist_ok=True)
    if os.path.exists(gz_path):
        return gz_path
    t = time.time()
    tmp = gz_path + f".tmp{os.getpid()}"
    gzip = [os.path.join(ROOT, "tools", "_build", "ppgzip"), "-l", "6", "-", tmp]
    if reads <= SHARD_READS:
        p1 = subprocess.Popen(gen_cmd(reads, seed, fixed), stdout=subprocess.PIPE)
        p2 = subprocess.Popen(gzip, stdin=p1.stdout)
        p1.stdout.close()
        ok = p2.wait() == 0 and p1.wait() == 0
    else:
        shards, left, i = [], reads, 0
        while left > 0:
            n = min(left, SHARD_READS)
            fifo = os.path.join(d, f"shard{i}.fifo.{os.getpid()}")
            if os.path.exists(fifo):
                os.remove(fifo)
            os.mkfifo(fifo)
            shards.append((fifo, n, seed + i))
            left -= n
            i += 1
        gens = [subprocess.Popen(["sh", "-c", " ".join(gen_cmd(n, s, fixed)) + f" > {fifo}"]) for fifo, n, s in shards]
        cat = subprocess.Popen(["cat"] + [f for f, _, _ in shards], stdout=subprocess.PIPE)
        p2 = subprocess.Popen(gzip, stdin=cat.stdout)
        cat.stdout.close()
        ok = p2.wait() == 0 and cat.wait() == 0 and all(g.wait() == 0 for g in gens)
        for f, _, _ in shards:
            os.remove(f)
    if not ok:
        raise RuntimeError("corpus generation failed")
    os.replace(tmp, gz_path)
    log(f"[bench] corpus {reads} reads -> {os.path.getsize(gz_path)/1e6:.1f} MB gz in {time.time()-t:.1f}s")
    return gz_path


def lift_cap():
    # records longer than 32 768 B make the reference's CreateIndex throw (quirk H2): lifted only for
    # uncapped long reads, a documented extension
    return bool(LONG["mean"]) and not LONG["cap"]


def workload_text(args, file_reads):
    lens = (f"lognormal lengths (mean {LONG['mean']:g} bp, sigma {LONG['sigma']:g}"
            f"{', capped at ' + str(LONG['cap']) if LONG['cap'] else ', uncapped: record cap lifted'})"
            if LONG["mean"] else f"{args.fixed_len}bp")
    shards = "" if file_reads <= SHARD_READS else f" (shards of {SHARD_READS} reads, seeds {args.seed}..)"
    return (f"Generator seed {args.seed}, {file_reads} reads x {lens} single-end{shards}, gzip -6 "
            f"(one member), chunk {args.chunk}")


def config_dict(args, world, file_reads, chunks, records, ubytes, cbytes):
    """The SAME dictionary in both arms (every value is a property of the workload, not of the arm)."""
    return {"workload": workload_text(args, file_reads), "file_reads": file_reads, "chunks": chunks,
            "records": records, "uncompressed_bytes": ubytes, "compressed_bytes": cbytes,
            "partition": ("whole file on one GPU" if world == 1 else
                          f"one file, {world} contiguous chunk ranges of near-equal compressed size, one per GPU; "
                          f"each rank moves only its own byte range and windows"),
            "l2": f"inputs larger than L2 ({cbytes / world / 1e9:.2f} GB compressed in, "
                  f"{ubytes / world / 1e9:.2f} GB inflated out per GPU and step vs 126 MB L2)"}


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


def gpu_numa_node(gpu):
    """NUMA node of the GPU's PCIe root as the (possibly virtual) machine reports it; None if unknown."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(gpu)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return None


def host_numa_nodes():
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except Exception:
        return None


def bind_memory_to_node(node):
    """set_mempolicy(MPOL_PREFERRED, node) for this process, so the pinned buffers allocated next are
    placed on the GPU's own NUMA node (needs no CPU on that node).  Returns True when applied."""
    if node is None or node < 0:
        return False
    try:
        libc = C.CDLL(None, use_errno=True)
        mask = C.c_ulong(1 << node)
        rc = libc.syscall(238, 1, C.byref(mask), C.c_ulong(64))  # x86-64 __NR_set_mempolicy, MPOL_PREFERRED
        return rc == 0
    except Exception:
        return False


def oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    return O


def reference_arm(args, world, file_reads, cores):
    """The host restatement of the reference's parallel DecompressAll, all cores.  Nothing of the
    product is imported: corpus tools + oracle only."""
    make_native(["tools", "oracle"])
    O = oracle()
    gz_path = make_gz(file_reads, args.fixed_len, args.seed, args.chunk)
    idx_path = corpus_paths(file_reads, args.fixed_len, args.seed, args.chunk)[2]
    gz = np.fromfile(gz_path, np.uint8)
    if not os.path.exists(idx_path):
        t = time.time()
        ox = O.OracleIndex.build(gz, args.chunk, lift_cap())
        tmp = idx_path + f".tmp{os.getpid()}"
        ox.serialize(tmp)          # IndexIO v0 bytes: the same file the product's CreateIndex writes
        os.replace(tmp, idx_path)
        log(f"[bench] oracle CreateIndex chunk={args.chunk}: {ox.count} points in {time.time()-t:.1f}s")
    ox = O.OracleIndex.deserialize(idx_path)
    times, recs, nbytes = [], 0, 0
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        recs, nbytes = O.decompress_all_mt(gz, ox, threads=cores)
        if recs < 0:
            raise RuntimeError(f"oracle failed rc={recs}")
        if i >= args.warmup:
            times.append(time.perf_counter() - t)
    dt = float(np.mean(times))
    ins = ox.inputs()
    cbytes = int(min(ins[-1], gz.size) - ((max(ins[0] - 1, 0)) & ~127))
    val = nbytes / dt / 1e9
    line = {
        "impl": "reference", "metric": "DecompressAll uncompressed GB/s", "value": val, "unit": "GB/s",
        "reads_per_s": recs / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": config_dict(args, world, file_reads, ox.count - 1, recs, nbytes, cbytes),
        "note": "C restatement of the reference's thread-pool DecompressAll (oracle/pp_oracle.c) on the same system "
                "zlib, all host cores, the whole file every step; the C# reference cannot run in this image (no dotnet)",
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": "whole workload, every step"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads PER GPU (the file holds gpus x reads)")
    ap.add_argument("--chunk", type=int, default=10_000)
    ap.add_argument("--fixed-len", type=int, default=150)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-oracle-gate", action="store_true", help="skip the per-chunk oracle digest gate (N = 1)")
    ap.add_argument("--no-create-index", action="store_true", help="skip the GPU CreateIndex leg (N = 1)")
    ap.add_argument("--lognormal", type=float, nargs=2, metavar=("MEAN", "SIGMA"),
                    help="read lengths lognormal with this mean (BASELINE config 4: --lognormal 10000 0.5 --chunk 1000)")
    ap.add_argument("--cap", type=int, default=0, help="clamp read lengths (16000 keeps records reference-legal, H2)")
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.lognormal:
        LONG.update(mean=args.lognormal[0], sigma=args.lognormal[1], cap=args.cap)
        args.fixed_len = 0
    file_reads = args.reads * max(world, args.gpus if args.impl == "reference" else 1)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        return reference_arm(args, max(world, args.gpus), file_reads, cores) if rank == 0 else 0

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

    def sync():
        torch.cuda.synchronize()

    if local_rank == 0:
        make_native(["tools", "oracle", os.path.join("parallelparsing_b200", "csrc")])
        make_gz(file_reads, args.fixed_len, args.seed, args.chunk)
    barrier()
    import parallelparsing_b200 as pp
    L = pp.lib()
    _, gz_path, idx_path = corpus_paths(file_reads, args.fixed_len, args.seed, args.chunk)
    if local_rank == 0 and not os.path.exists(idx_path):
        t = time.time()
        # corpus preparation, untimed: the index is built on the GPU (pp_index_create_gpu, byte-identical to the host
        # pass — tests/test_gpu_parity.py, and checked again in this run's create_index leg at N = 1); the host
        # pass (11 s per 10 M reads) is the fallback for what the GPU entry point declines
        how = "GPU"
        try:
            ix0 = pp.Core.BuildDeflateIndexGpu(gz_path, args.chunk, pp.Device(local_rank), lift_record_cap=lift_cap())
        except pp.ZException as e:
            how = f"host (GPU build declined: {e.Code})"
            ix0 = pp.Core.BuildDeflateIndex(gz_path, args.chunk, lift_record_cap=lift_cap())
        tmp = idx_path + f".tmp{os.getpid()}"
        pp.IndexIO.Serialize(ix0, tmp)
        os.replace(tmp, idx_path)
        log(f"[bench] CreateIndex on the {how}, chunk={args.chunk}: {ix0.Count} points in {time.time()-t:.1f}s")
        del ix0
    barrier()

    # pinned host memory on the GPU's own NUMA node where the machine exposes one
    node = gpu_numa_node(local_rank)
    bound = bind_memory_to_node(node) if (host_numa_nodes() or 1) > 1 else False
    ix = pp.IndexIO.Deserialize(idx_path)
    gz_len = os.path.getsize(gz_path)
    first, n_mine = pp.partition_chunks(ix, world)[rank]
    dev = pp.Device(local_rank)
    # compact windows (zlib-compressed over PCIe, a GPU pre-pass unpacks them) pay when the windows are a
    # visible share of the bytes moved: 3 % at chunk 10,000 (the pre-pass costs more than it saves), 27 % at chunk 1,000
    compact = 32768.0 * (ix.Count - 1) > 0.08 * gz_len
    # at most two jobs are alive at a time (a job holds the slots of its whole partition: 39 GB at 100 M reads)
    job = pp.Job(dev, ix, gz_len, first, n_mine)                                                 # staged, plain: `value`
    lo, ln = job.file_range()
    rng_ptr = C.c_void_p()
    pp.check(L.pp_host_alloc(max(ln, 1), C.byref(rng_ptr)), "pp_host_alloc")
    rng = np.ctypeslib.as_array(C.cast(rng_ptr, C.POINTER(C.c_uint8)), shape=(max(ln, 1),))
    with open(gz_path, "rb") as f:    # this rank's byte range only
        f.seek(lo)
        got = f.readinto(memoryview(rng)[:ln])
        assert got == ln, (got, ln)

    def run(j, to_host=None):
        j.upload_range(rng_ptr, lo, ln)
        if to_host is None:
            j.execute()
        else:
            j.execute_to_host(*to_host)
        j.download()
        return j.info()

    # ---------------------------------------------------------------- correctness gate before timing
    info = run(job)
    if info.status != 0:
        raise SystemExit(f"bench.py: DecompressAll failed status={info.status}")
    U, R, Us, n_chunks = info.total_bytes, info.total_records, info.scanned_bytes, info.n_chunks
    h2d_plain = info.h2d_bytes
    outs = ix.scalars()[0]
    if U != int(outs[first + n_mine] - outs[first]):
        raise SystemExit(f"bench.py: rank {rank} produced {U} bytes, expected {int(outs[first + n_mine] - outs[first])}")
    bd, fd = job.digests()
    sig = [(job.chunk(k).inflated, job.chunk(k).records) for k in range(n_chunks)]
    def check_mode(name, j):
        i2 = run(j)
        b2, f2 = j.digests()
        if i2.status != 0 or (i2.total_bytes, i2.total_records) != (U, R) or not (np.array_equal(bd, b2) and np.array_equal(fd, f2)):
            raise SystemExit(f"bench.py: {name} DecompressAll disagrees with the staged run")
    gate = "modes agree chunk by chunk (digests of bytes and of record fields)"
    if world == 1 and not args.no_oracle_gate:
        # every chunk against the oracle: boundaries (its own CreateIndex), length, records, bytes digest,
        # fields digest — and the gzip trailer's CRC-32 / ISIZE over the streamed-to-host bytes below
        O = oracle()
        gz_np = np.fromfile(gz_path, np.uint8)
        t = time.time()
        ox = O.OracleIndex.build(gz_np, args.chunk, lift_cap())
        if ox.inputs() != [int(x) for x in ix.scalars()[1]]:
            raise SystemExit("bench.py: chunk boundaries differ from the oracle's CreateIndex")
        want = O.chunk_digests(gz_np, ox)
        for k in range(n_chunks):
            if (sig[k][0], sig[k][1], int(bd[k]), int(fd[k])) != tuple(int(x) for x in want[k]):
                raise SystemExit(f"bench.py: chunk {k} differs from the oracle")
        gate = f"every chunk == oracle (length, records, bytes digest, fields digest; {time.time()-t:.0f}s)"
        del want
    if R < args.reads * (n_mine > 0) and world == 1:
        raise SystemExit(f"bench.py: {R} records < {args.reads} reads")
    log(f"[bench] rank {rank}: chunks [{first}, {first + n_mine}), {U/1e9:.2f} GB, gate: {gate}")

    # pinned destinations for the to-host variants
    out_ptr = C.c_void_p()
    to_host_leg = U <= 16_000_000_000    # the whole inflated stream in pinned host memory: skipped for very large partitions
    if to_host_leg:
        pp.check(L.pp_host_alloc(max(U, 1), C.byref(out_ptr)), "pp_host_alloc")
    lines_ptr = C.c_void_p()
    pp.check(L.pp_host_alloc(max(16 * R, 16), C.byref(lines_ptr)), "pp_host_alloc")
    lp = [C.c_void_p(lines_ptr.value + 4 * R * f) for f in range(4)]

    sampler = ClockSampler(local_rank)
    sampler.start()

    def timed(step, steps, warmup):
        """warmup, barrier+sync, K steps, sync (the library's streams are joined by the device-wide sync)."""
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
    job.upload_range(rng_ptr, lo, ln)
    sync()
    t_dev = timed(job.execute, args.steps, args.warmup)
    job.download()
    info = job.info()
    t_inflate, t_parse, t_scan = info.inflate_ms * 1e-3, info.parse_ms * 1e-3, info.scan_ms * 1e-3
    launches_per_step = info.launches

    # --- H2D link probe: this rank's range + windows by cudaMemcpyAsync, all ranks at once ---------
    def probe():
        job.upload_range(rng_ptr, lo, ln)
    t_probe = timed(probe, 3, 1)
    link_gbs = h2d_plain / t_probe / 1e9 if t_probe > 0 else 0.0

    job.free()   # its device memory goes back to the pool for the two end-to-end jobs

    # --- end to end through the C ABI, host buffers every step ------------------------------------
    e2e_steps = max(3, args.steps)
    job_pipe = pp.Job(dev, ix, gz_len, first, n_mine, pipeline=True, compact_windows=compact)    # e2e: pipelined
    check_mode("pipelined", job_pipe)
    t_pipe = timed(lambda: run(job_pipe), e2e_steps, 2)
    h2d_bytes, d2h_bytes = job_pipe.info().h2d_bytes, job_pipe.info().d2h_bytes
    job_zc = pp.Job(dev, ix, gz_len, first, n_mine, zero_copy=True, compact_windows=compact)     # e2e: pull
    check_mode("pull", job_zc)
    t_pull = timed(lambda: run(job_zc), e2e_steps, 2)
    best = job_pipe if t_pipe <= t_pull else job_zc
    (job_zc if best is job_pipe else job_pipe).free()

    def step_offsets():
        run(best)
        pp.check(L.pp_job_fetch_line_starts(best.h, *lp), "fetch_line_starts")
    t_offsets = timed(step_offsets, e2e_steps, 1)

    def step_bytes():
        run(best, to_host=(out_ptr, max(U, 1)))
        pp.check(L.pp_job_fetch_line_starts(best.h, *lp), "fetch_line_starts")
    t_bytes = timed(step_bytes, max(2, e2e_steps // 2), 1) if to_host_leg else float("inf")
    crc_ok = None
    if world == 1 and to_host_leg:
        import zlib
        host_bytes = np.ctypeslib.as_array(C.cast(out_ptr, C.POINTER(C.c_uint8)), shape=(max(U, 1),))[:U]
        with open(gz_path, "rb") as f:
            f.seek(-8, 2)
            tr = np.frombuffer(f.read(8), "<u4")
        crc_ok = bool((zlib.crc32(memoryview(host_bytes)) & 0xffffffff) == int(tr[0]) and (U & 0xffffffff) == int(tr[1]))
        if not crc_ok:
            raise SystemExit("bench.py: CRC-32/ISIZE of the bytes streamed to the host != gzip trailer")

    flags_best = (4 if best is job_pipe else 2) | (8 if compact else 0)

    def step_one_call():
        h = C.c_void_p()
        base = C.c_void_p(rng_ptr.value - lo)   # the job addresses the file as gz[offset]; only our range is touched
        rc = L.pp_decompress_all(dev.h, ix.h, base, gz_len, first, n_mine, flags_best, C.byref(h))
        if rc != 0:
            raise SystemExit(f"bench.py: pp_decompress_all rc={rc}")
        L.pp_job_free(h)
    t_one = timed(step_one_call, max(2, e2e_steps // 2), 1)
    barrier()
    clocks = sampler.finish()

    # sums and maxima over ranks
    vals = torch.tensor([t_dev, t_pipe, t_pull, t_inflate, t_parse, t_offsets, t_bytes, t_one, t_probe],
                        device="cuda", dtype=torch.float64)
    sums = torch.tensor([U, R, Us, n_chunks, info.compressed_bytes, h2d_bytes, d2h_bytes, h2d_plain], device="cuda",
                        dtype=torch.float64)
    links = torch.tensor([link_gbs], device="cuda", dtype=torch.float64)
    links_all = [torch.zeros_like(links) for _ in range(world)]
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_gather(links_all, links)
    else:
        links_all = [links]
    t_dev, t_pipe, t_pull, t_inflate_max, t_parse_max, t_offsets, t_bytes, t_one, t_probe = [float(x) for x in vals.tolist()]
    Ut, Rt, Ust, chunks_t, comp_t, h2d_t, d2h_t, h2d_plain_t = [int(x) for x in sums.tolist()]
    per_gpu_link = [float(x.item()) for x in links_all]

    if rank == 0:
        # compressed bytes of the WHOLE file's chunk list, as the reference arm computes them (the per-rank
        # ranges overlap by their alignment padding, so their sum is a few hundred bytes more)
        ins_all = ix.scalars()[1]
        comp_file = int(min(int(ins_all[-1]), gz_len) - ((max(int(ins_all[0]) - 1, 0)) & ~127))
        peak, peak_src = load_peaks()
        b_parse = Us + 16 * R                      # this rank's launch: bytes scanned + four u32 line starts / record
        b_inflate = info.compressed_bytes + 32768 * n_chunks + U
        traffic = load_ncu_traffic(args.reads, args.chunk) if world == 1 else {}
        t_e2e = min(t_pipe, t_pull)
        moved = h2d_t if t_pipe <= t_pull else h2d_t   # bytes that cross PCIe per step (same in both modes: compact windows)
        line = {
            "metric": "DecompressAll uncompressed GB/s", "value": Ut / t_dev / 1e9, "unit": "GB/s",
            "reads_per_s": Rt / t_dev,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args, world, file_reads, chunks_t, Rt, Ut, comp_file),
            "gate": gate + ("; CRC-32/ISIZE of the bytes streamed to the host == gzip trailer" if crc_ok else ""),
            "e2e": {"value": Ut / t_e2e / 1e9, "unit": "GB/s", "reads_per_s": Rt / t_e2e,
                    "ms_per_step": t_e2e * 1e3, "h2d_bytes_per_step": h2d_t, "d2h_bytes_per_step": d2h_t,
                    "mode": "pipelined" if t_pipe <= t_pull else "pull",
                    "pipelined": {"value": Ut / t_pipe / 1e9, "ms_per_step": t_pipe * 1e3,
                                  "mode": "hybrid: the first wave of chunks (one per resident CTA) is pulled by the kernel "
                                          "from pinned host memory, the rest goes by cudaMemcpyAsync in 8 MB pieces on a copy "
                                          "stream (held back until the first wave is nearly through its input); chunks wait on "
                                          "a device-side byte counter"},
                    "pull": {"value": Ut / t_pull / 1e9, "ms_per_step": t_pull * 1e3,
                             "mode": "kernels read the compressed range from pinned host memory (TMA over PCIe)"},
                    "windows": (f"zlib-compressed over PCIe, inflated on the GPU (PP_JOB_COMPACT_WINDOWS): {h2d_t} B moved per "
                                f"step vs {h2d_plain_t} B with raw 32 KB windows") if compact else
                               "raw 32 KB windows (3 % of the bytes moved at this chunk size: compressing them does not pay)",
                    "one_call": {"value": Ut / t_one / 1e9, "ms_per_step": t_one * 1e3,
                                 "what": "cold pp_decompress_all + pp_job_free every step (plan, allocations, pinning included)"},
                    "with_line_offsets_to_host": {"value": Ut / t_offsets / 1e9, "ms_per_step": t_offsets * 1e3,
                                                  "d2h_bytes_per_step": d2h_t + 16 * Rt, "into": "pinned memory, every step"},
                    "with_bytes_to_host": {"value": Ut / t_bytes / 1e9 if t_bytes != float("inf") else None,
                                           "ms_per_step": t_bytes * 1e3 if t_bytes != float("inf") else None,
                                           "d2h_bytes_per_step": d2h_t + 16 * Rt + Ut,
                                           "what": "the whole inflated stream + line offsets into pinned host memory, the "
                                                   "D2H of each chunk queued as soon as the kernel flags it (overlaps the decode)"},
                    "host_link": {"h2d_probe_gbs_per_gpu": per_gpu_link, "h2d_probe_gbs_total": h2d_plain_t / t_probe / 1e9,
                                  "e2e_h2d_gbs_total": moved / t_e2e / 1e9,
                                  "what": "cudaMemcpyAsync of every rank's range + windows, all ranks at once, vs the "
                                          "H2D rate the e2e step sustains",
                                  "numa": {"host_nodes": host_numa_nodes(), "gpu_node_rank0": node,
                                           "mempolicy_applied_rank0": bound}}},
            "roofline": {"kernel": "pp_parse_kernel", "bound": "hbm", "achieved": b_parse / t_parse / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": b_parse / t_parse / 1e9 / peak,
                         "traffic": traffic.get("pp_parse_kernel"),
                         "algorithmic_bytes": b_parse, "ms": t_parse * 1e3, "peak_source": peak_src,
                         "timing": "CUDA events on the library stream around the kernel, last timed step, rank 0",
                         "peak_kind": "device COPY bandwidth (read + write); this kernel reads ~25x more than it "
                                      "writes, and a read-only stream can run above that figure"},
            "inflate": {"kernel": "pp_inflate_kernel",
                        "bound": "instruction issue / shared-memory latency (Huffman decode + LZ77 resolve), not HBM",
                        "decompressed_gbs": U / t_inflate / 1e9, "ms": t_inflate * 1e3,
                        "algorithmic_bytes": b_inflate, "hbm_frac": b_inflate / t_inflate / 1e9 / peak,
                        "traffic": traffic.get("pp_inflate_kernel"),
                        "share_of_step": t_inflate / t_dev},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if world == 1 and U <= (16 << 30) and not args.no_create_index:
            # CreateIndex on the GPU (pp_index_create_gpu) for the same file: wall clock from pinned host memory,
            # everything included; the result must serialize to the very index file this run decoded with
            # (built by the host pass, pp_index_create = Core.BuildDeflateIndex on zlib, one thread)
            gz_all, gz_all_ptr = pp.pinned_copy(np.fromfile(gz_path, np.uint8))
            ci = []
            for _ in range(3):
                t0 = time.perf_counter()
                gix, st = pp.Core.BuildDeflateIndexGpu(gz_all, args.chunk, dev, lift_record_cap=lift_cap(), want_stats=True)
                ci.append(((time.perf_counter() - t0) * 1e3, st))
            L.pp_host_free(gz_all_ptr)
            ms, st = min(ci, key=lambda x: x[0])
            t0 = time.perf_counter()
            hix = pp.Core.BuildDeflateIndex(gz_path, args.chunk, lift_record_cap=lift_cap())   # the host pass: zlib, one thread
            host_s = time.perf_counter() - t0
            tg, th = idx_path + f".gpu{os.getpid()}", idx_path + f".host{os.getpid()}"
            pp.IndexIO.Serialize(gix, tg)
            pp.IndexIO.Serialize(hix, th)
            with open(tg, "rb") as fa, open(th, "rb") as fb, open(idx_path, "rb") as fc:
                ga = fa.read()
                same = ga == fb.read() and ga == fc.read()
            os.remove(tg)
            os.remove(th)
            if not same:
                raise SystemExit("bench.py: the index built on the GPU differs from the host-built index")
            line["create_index"] = {"value": U / ms / 1e6, "unit": "GB/s", "ms": ms, "points": st["points"],
                                    "blocks": st["blocks"], "identical_to_host_built_index": same,
                                    "stages_ms": {k: round(v, 3) for k, v in st.items() if k.endswith("_ms")},
                                    "host_zlib_1thread": {"value": U / host_s / 1e9, "unit": "GB/s", "s": host_s},
                                    "what": "pp_index_create_gpu, host buffer in, index out (best of 3 cold calls)"}
        if not args.no_cpu_baseline and world == 1:
            O = oracle()
            gz_np = np.fromfile(gz_path, np.uint8)
            ox = O.OracleIndex.deserialize(idx_path)
            ts = []
            for i in range(3):
                t0 = time.perf_counter()
                recs, nbytes = O.decompress_all_mt(gz_np, ox, threads=cores)
                if i:
                    ts.append(time.perf_counter() - t0)
            dt = float(np.mean(ts))
            assert recs == R and nbytes == U, (recs, R, nbytes, U)
            line["cpu_baseline"] = {"value": nbytes / dt / 1e9, "unit": "GB/s", "reads_per_s": recs / dt,
                                    "cores": cores, "kind": "port",
                                    "sample": "whole workload, 2 timed passes after 1 warm-up (oracle thread pool)"}
            # the reference's serial baseline (SimpleDecompressor: GZipStream + line parser), one thread,
            # on a bounded prefix of the same file (a truncated stream ends in an error: only the bytes
            # it got through are used)
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
    best.free()
    for p in (rng_ptr, out_ptr, lines_ptr):
        if p:
            L.pp_host_free(p)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""bench.py -- headline benchmark: GiB/s of AES-128-CTR over a 16 GiB device-resident buffer per
B200 (BASELINE.json `metric`), as absolute throughput and as a fraction of the
measured HBM roofline, next to micro_aes.c timed on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W]             (N=1: plain python)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        the reference's own CPU code on all host cores

One step = one pass of the hot path (one ctr_queue8_kernel launch: table-driven warps plus the bitsliced
ALU co-runner warps sharing the range through a work queue) over this rank's 16 GiB shard of the
N*16 GiB buffer; rank r owns keystream blocks [r*2^30, (r+1)*2^30) (counter-range sharding, no
data-path collective; the key and IV are broadcast once over NCCL).  Scaling is therefore weak.
torch is used for device memory, streams/events and torch.distributed only; the encryption is
libuaes_b200.so called through its C ABI.

After the timed headline the default run also measures BASELINE configs 2-4 (`secondary`: CTR 1 GiB,
XTS-256 with 512-byte sectors, GCM-128 4 GiB + tag -- at N > 1 sharded like the headline, GCM with its
one 16-byte all-gather and the combine inside every step), checks windows of every rank's output
against the oracle (all-reduced), and runs the e2e legs on host buffers.  Other workloads:
--workload xts256 | gcm128 | ... (README.md).
"""
import argparse
import ctypes
import importlib
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30
KEY = bytes.fromhex("279fb74a7572135e8f9b8ef6d1eee003")        # main.c:19 cipherKey[0..16)
IV = bytes.fromhex("8ea2b7ca516745bfeafc4990")                 # main.c:18 iVec[0..12)
SEED = 0x5EED0002
METRIC = "AES-128-CTR throughput, 16 GiB buffer per B200"
UNIT = "GiB/s"


def shard_plan(total_blocks, world):
    """contiguous keystream-block ranges per rank: [(first_block, nblocks)]"""
    per = total_blocks // world
    return [(r * per, per if r < world - 1 else total_blocks - r * per) for r in range(world)]


def gcm_shards(total_bytes, world):
    """byte ranges [(offset, nbytes)] of one GCM message over `world` ranks: whole 16-byte blocks
    per rank, the last rank owns the ragged tail"""
    blocks = total_bytes // 16
    plan = shard_plan(blocks, world)
    out = [(f * 16, n * 16) for f, n in plan]
    off, n = out[-1]
    out[-1] = (off, total_bytes - off)
    return out


def gcm_blocks_after(shards, total_bytes):
    """GHASH blocks after the end of each shard (the exponents of H in the combine step)"""
    total_blocks = (total_bytes + 15) // 16
    return [total_blocks - (off + n + 15) // 16 for off, n in shards]


# --------------------------------------------------------------------------- CPU baseline

def _ref_lib():
    """(path, kind): the compiled unmodified reference if it travelled with the repo, else the
    oracle port"""
    p = os.path.join(ROOT, "oracle", "_ref", "libref128.so")
    if os.path.exists(p):
        return p, "reference"
    q = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(q):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return q, "port"


def _cpu_worker(args):
    """one forked process = one core: AES_CTR_encrypt over its own slice, timed around the call
    (the reference keeps its round keys in a file-static array, micro_aes.c:72, so threads
    would race: processes, as BASELINE.md section 4 prescribes)"""
    path, kind, nbytes, reps = args
    lib = ctypes.CDLL(path)
    src = ctypes.create_string_buffer(os.urandom(4096) * (nbytes // 4096), nbytes)
    dst = ctypes.create_string_buffer(nbytes)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        if kind == "reference":
            lib.AES_CTR_encrypt(KEY, IV, src, ctypes.c_size_t(nbytes), dst)
        else:
            lib.oracle_ctr_crypt(128, KEY, IV, src, ctypes.c_size_t(nbytes), dst)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best


def cpu_throughput(slice_mib=64, reps=1, cores=None):
    """aggregate GiB/s of the CPU implementation with one process per host core"""
    path, kind = _ref_lib()
    cores = cores or os.cpu_count() or 1
    nbytes = slice_mib << 20
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(path, kind, nbytes, reps)] * cores)
        wall = time.perf_counter() - t0
    # every process ran concurrently on its own slice: aggregate = total bytes / slowest process
    agg = cores * nbytes / max(times) / GIB
    return {"value": round(agg, 4), "unit": UNIT, "cores": cores, "kind": kind,
            "per_core_MiB_s": round(nbytes / statistics.median(times) / (1 << 20), 2),
            "sample": f"{cores} forked processes x {slice_mib} MiB AES-128-CTR "
                      f"({'micro_aes.c, gcc -O2 -fno-strict-aliasing' if kind == 'reference' else 'oracle/aes_oracle.c'}),"
                      f" best of {reps}, wall {wall:.1f}s"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample per step: cores x 32 MiB (~1 s per step at ~32 MiB/s/core)
    for _ in range(args.warmup):
        cpu_throughput(slice_mib=8, cores=cores)
    vals, t0 = [], time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_throughput(slice_mib=32, cores=cores)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = statistics.median(vals)
    cb = dict(last, value=round(v, 4))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1000 * wall / max(args.steps, 1), 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "AES-128-CTR, 16 GiB buffer per B200 (reference arm: bounded sample "
                               f"of {cores} x 32 MiB per step on {cores} host cores)"},
        "cpu_baseline": cb,
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------- GPU arm

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        out = self.p.communicate()[0]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def mem_available_gib():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / (1 << 20)
    except OSError:
        pass
    return 0.0


KERNEL_NAMES = {
    "ctr128": "uaes::ctr_queue8_kernel<10,512,256,1,64> (16 table-driven warps, one row in flight, 64 registers + 8 narrow bitsliced warps, 112 registers; two-ended work queue)",
    "ctr256": "uaes::ctr_queue8_kernel<14,384,256,2,0> (12 table-driven + 8 narrow bitsliced warps, 96 registers each)",
    "ecb128": "uaes::ecb_hybrid_kernel<10,false> (table-driven + bitsliced warps)",
    "ecb128dec": "uaes::ecb_dec_hybrid_kernel<10,false> (16 table-driven warps + 4 bitsliced inverse-cipher warps, work queue)",
    "xts256": "uaes::xts_sectors_hybrid_kernel<14,true> (16 table-driven warps, one sector in flight + 4 bitsliced warps, work queue)",
    "xts256dec": "uaes::xts_sectors_hybrid_kernel<14,false> (16 table-driven warps + 4 bitsliced inverse-cipher warps, work queue)",
    "xts256unit": "uaes::xts_unit_hybrid_kernel<14> (16 table-driven warps, one row in flight + 4 bitsliced warps, static split)",
    "gcm128": "uaes::gcm_setup_kernel + uaes::gcm_bulk_hybrid_kernel<10,0> (table-driven + bitsliced warps) + uaes::gcm_finish_kernel",
    "gcmsiv128": "uaes::gcm_bulk_kernel<10,1,true> (POLYVAL) + uaes::ctr32_kernel<10>",
    "ocb128": "uaes::ocb_hybrid_kernel<10> (table-driven + bitsliced warps)",
    "cbc128dec": "uaes::ecb_dec_hybrid_kernel<10,true> (CBC form; 16 table-driven warps + 4 bitsliced inverse-cipher warps, work queue)",
    "cfb128dec": "uaes::ecb_hybrid_kernel<10,true> (CFB form)",
    "ccm128batch": "uaes::ccm_batch_kernel<10> (1 KiB messages, one per lane)",
    "eax128batch": "uaes::eax_batch_kernel<10> (1 KiB messages, one per lane)",
    "siv128batch": "uaes::siv_batch_kernel<10> (1 KiB messages, one per lane)",
    "gcm128batch": "uaes::gcm_batch_kernel<10> (1 KiB messages, one per lane)",
}
LOOKUPS_PER_BLOCK = {"ctr128": 128, "ctr256": 192}          # table-driven warps, after the round-1/2 hoisting
SMEM_LOOKUP_PEAK = 31.4                                     # conflict-free LDS.32 lane-lookups / clk / SM, measured (profiles/r1_microbench.log)
N_SM = 148


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    uaes = importlib.import_module("micro-aes_b200")
    core = uaes.core()            # raises if the CUDA library is not built: no fallback

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- the one collective of the CTR / XTS path: broadcast key || iv from rank 0
    kiv = torch.from_numpy(np.frombuffer(KEY + IV if rank == 0 else bytes(28), dtype=np.uint8).copy()).cuda()
    if world > 1:
        dist.broadcast(kiv, src=0)
    kb = bytes(kiv.cpu().numpy())
    key, iv = kb[:16], kb[16:]
    key32 = key + bytes(range(16))
    keys64 = key32 + bytes(range(32, 64))

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        from util import Oracle
        orc = Oracle()
    except Exception as e:                                     # checker missing: report, do not hide
        orc, orc_err = None, str(e)

    nbytes = int(args.gib_per_gpu * GIB)
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst = torch.empty(nbytes + 16, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    uaes.set_stream(stream.cuda_stream)

    def all_true(flag):
        """a parity flag counts only if EVERY rank saw it (rank 3 owns the 2^32 carry)"""
        if world == 1 or not isinstance(flag, bool):
            return flag
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warmup, sample_clocks=False):
        uaes.set_async(True)
        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local) if sample_clocks and rank == 0 else None
        launches0 = uaes.kernel_launches()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record(stream)
        for i in range(steps):
            step()
            evs[i + 1].record(stream)
        barrier()
        launches = uaes.kernel_launches() - launches0
        clocks = sampler.stop() if sampler else None
        uaes.set_async(False)
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        total = evs[0].elapsed_time(evs[-1])
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return per, float(t.item()), int(launches), clocks

    # ------------------------------------------------------------------ workloads
    class Work:
        """one workload on this rank's shard: step() + check() -> True / False / 'unchecked: why'"""
        def __init__(self, wl, n):
            self.wl, self.n = wl, n
            self.first_block = shard_plan((n // 16) * world, world)[rank][0]
            self.seed = SEED
            self.extra = {}
            uaes.fill_splitmix64(self.seed, self.first_block * 2, src, n // 8)
            if wl.endswith("batch"):
                self.msg_bytes = 1024
                self.nmsg = n // self.msg_bytes
                rec = np.zeros(self.nmsg, dtype=np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("aad_off", "<u8"), ("len", "<u4"),
                                                         ("aad_len", "<u4"), ("nonce", "u1", 16), ("result", "<i4"), ("reserved", "<u4")]))
                idx = np.arange(self.nmsg, dtype=np.uint64)
                rec["in_off"], rec["out_off"], rec["aad_off"] = idx * self.msg_bytes, idx * (self.msg_bytes + 16), (idx % 4096) * 16
                rec["len"], rec["aad_len"] = self.msg_bytes, 16
                rec["nonce"][:, :8] = idx.view(np.uint8).reshape(-1, 8)
                self.msgs_dev = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).cuda()
                self.bdst = torch.empty(self.nmsg * (self.msg_bytes + 16), dtype=torch.uint8, device="cuda")
            if wl == "gcm128" and world > 1:
                self.mine = torch.zeros(16, dtype=torch.uint8, device="cuda")
                self.allp = torch.zeros(16 * world, dtype=torch.uint8, device="cuda")
                shards = gcm_shards(n * world, world)
                self.after = gcm_blocks_after(shards, n * world)
                self.tag, self.tag_dev = None, torch.zeros(16, dtype=torch.uint8, device="cuda")

        def step(self):
            wl, n, fb = self.wl, self.n, self.first_block
            if wl.endswith("batch"):
                uaes.ccm_batch(128, key32 if wl[:3] == "siv" else key, self.msgs_dev, self.nmsg, src[:65536], src, self.bdst, mode=wl[:3])
            elif wl == "ctr128":
                uaes.ctr_crypt_range(128, key, iv, fb, src, n, dst)
            elif wl == "ctr256":
                uaes.ctr_crypt_range(256, key32, iv, fb, src, n, dst)
            elif wl == "ecb128":
                uaes.ecb(128, key, src, n, dst, True)
            elif wl == "ecb128dec":
                uaes.ecb(128, key, src, n, dst, False)
            elif wl == "xts256unit":    # the reference's AES_XTS_encrypt: ONE data unit of world * n bytes, this rank's block range
                uaes.xts_crypt_range(256, keys64, IV + bytes(4), fb, src, n, dst, True)
            elif wl == "xts256dec":
                uaes.xts_sectors(256, keys64, fb // 32, 512, src, n, dst, False)
            elif wl == "ocb128":
                uaes.ocb(128, key, iv, b"", src, n, dst, True)
            elif wl == "cbc128dec":
                uaes.chain_decrypt(128, key, key, src, n, dst, cbc=True)
            elif wl == "cfb128dec":
                uaes.chain_decrypt(128, key, key, src, n, dst, cbc=False)
            elif wl == "gcmsiv128":     # two passes (POLYVAL, then CTR)
                uaes.gcmsiv(128, key, iv, b"", src, n, dst, True)
            elif wl == "xts256":        # BASELINE config 3: 512-byte sectors, sector numbers follow the shard
                uaes.xts_sectors(256, keys64, fb // 32, 512, src, n, dst, True)
            elif wl == "gcm128" and world == 1:   # BASELINE config 4: one message, one GPU
                uaes.gcm_encrypt(128, key, iv, b"", src, n, dst)
            elif wl == "gcm128":
                # ONE message of world * n bytes sharded by block range: fused CTR+GHASH pass per rank, the
                # 16-byte contribution stays on the GPU, ONE device-to-device all-gather of 16 B per rank,
                # rank 0 folds the contributions into the tag
                uaes.gcm_shard(128, key, iv, fb, src, n, dst, partial_dev=self.mine)
                dist.all_gather_into_tensor(self.allp, self.mine)
                if rank == 0:      # the tag stays on the GPU: nothing in the step waits for the host
                    uaes.gcm_combine(128, key, iv, b"", None, self.after, n * world, partials_dev=self.allp, tag_dev=self.tag_dev)

        # ---- parity inside the bench: windows of this rank's shard against the oracle, on EVERY rank
        def check(self):
            if orc is None:
                return f"unchecked: {orc_err}"
            wl, n, fb = self.wl, self.n, self.first_block
            W = 65536
            offs = sorted({0, n - W, (n // 2) // W * W, (n // 3) // 4096 * 4096})
            pt = lambda off, ln=W: orc.splitmix(self.seed, fb * 2 + off // 8, ln // 8)
            got = lambda off, ln=W: bytes(dst[off:off + ln].cpu().numpy())
            if wl in ("ctr128", "ctr256"):
                k = key if wl == "ctr128" else key32
                # the 2^32 carry of the counter field, if this shard crosses it (rank 3 of 8 at 16 GiB per GPU)
                cross = ((1 << 32) - 1 - fb) * 16
                if 0 <= cross < n - W:
                    offs.append(cross // 16 * 16 - W // 2 if cross > W else 0)
                return all(got(o) == orc.ctr(k, iv, pt(o), first_block=fb + o // 16) for o in offs)
            if wl == "xts256":
                return all(got(o) == orc.xts_sectors(keys64, fb // 32 + o // 512, 512, pt(o))[1] for o in offs)
            if wl == "xts256unit":
                return all(got(o) == orc.xts_range(keys64, IV + bytes(4), fb + o // 16, pt(o))[1] for o in offs[:2])
            if wl == "ecb128":
                return all(got(o) == orc.ecb_encrypt(key, pt(o)) for o in offs)
            if wl == "gcm128":
                ok = all(got(o) == orc.ctr(key, iv, pt(o), first_block=1 + fb + o // 16) for o in offs)
                # GHASH against the oracle on the first MiB of the shard: a shard's contribution is the plain
                # absorb chain over its ciphertext (zero start state), wherever the shard sits in the message
                H = orc.encrypt_block(key, bytes(16))
                tmp = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
                z = uaes.gcm_shard(128, key, iv, fb, src, 1 << 20, tmp)
                ok &= z == orc.ghash_absorb(H, bytes(tmp.cpu().numpy()), 1 << 20) and bytes(tmp.cpu().numpy()) == got(0, 1 << 20)
                # the H-power algebra: the same range as 3 uneven shards must fold to the same tag
                cuts = [0, (n // 5) // 16 * 16, (n // 2 + 4096) // 16 * 16, n]
                parts, after = [], []
                big = torch.empty(max(b - a for a, b in zip(cuts, cuts[1:])), dtype=torch.uint8, device="cuda")
                for a, b in zip(cuts, cuts[1:]):
                    parts.append(uaes.gcm_shard(128, key, iv, fb + a // 16, src[a:], b - a, big))
                    after.append((n - b) // 16)
                del big
                t3 = uaes.gcm_combine(128, key, iv, b"", parts, after, n)
                if world == 1:
                    ok &= t3 == got(n, 16)
                    self.extra["tag"] = got(n, 16).hex()
                else:
                    whole = uaes.gcm_shard(128, key, iv, fb, src, n, dst)
                    ok &= t3 == uaes.gcm_combine(128, key, iv, b"", [whole], [0], n)
                    # and the combined tag of the timed loop (rank 0) == the tag folded from every rank's 3 sub-shards
                    sub = torch.from_numpy(np.frombuffer(b"".join(parts), dtype=np.uint8).copy()).cuda()
                    allsub = torch.zeros(48 * world, dtype=torch.uint8, device="cuda")
                    dist.all_gather_into_tensor(allsub, sub)
                    if rank == 0:
                        self.tag = bytes(self.tag_dev.cpu().numpy())
                        tot = (n * world) // 16
                        aft = [tot - (r * (n // 16) + b // 16) for r in range(world) for b in cuts[1:]]
                        t_all = uaes.gcm_combine(128, key, iv, b"", None, aft, n * world, partials_dev=allsub)
                        ok &= self.tag is not None and t_all == self.tag
                        self.extra["tag"] = self.tag.hex() if self.tag else None
                return bool(ok)
            return "unchecked: spot check covers ctr128/ctr256/ecb128/xts256/xts256unit/gcm128 (all modes: tests/)"

    def measure(wl, n, steps, warmup, sample_clocks=False):
        w = Work(wl, n)
        per, total_max, launches, clocks = timed(w.step, steps, warmup, sample_clocks)
        parity = all_true(w.check())
        kernel_ms = statistics.mean(per)
        return {"w": w, "per": per, "total_ms": total_max, "launches": launches, "clocks": clocks, "parity": parity,
                "kernel_ms": kernel_ms, "value": world * n * steps / GIB / (total_max / 1e3)}

    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak, peak_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass

    wl = args.workload
    head = measure(wl, nbytes, args.steps, args.warmup, sample_clocks=True)
    clocks = head["clocks"]

    # ---- which roof binds: the shared-memory lookup pipe of the table-driven warps
    lookup_roof, queue_split = None, None
    try:
        if wl in LOOKUPS_PER_BLOCK:
            tt, bs, ub = uaes.ctr_queue_stats()
            if tt + bs > 0:
                queue_split = {"table_driven_units": tt, "bitsliced_units": bs, "unit_blocks": ub,
                               "bitsliced_share": round(bs / (tt + bs), 4)}
                mhz = (clocks or {}).get("sm_mhz") or 1965.0
                lookups = tt * ub * LOOKUPS_PER_BLOCK[wl]
                ach = lookups / (head["kernel_ms"] / 1e3) / (mhz * 1e6) / N_SM
                lookup_roof = {"bound": "smem_lookup", "achieved": round(ach, 2), "peak": SMEM_LOOKUP_PEAK,
                               "unit": "lane-lookups/clk/SM", "frac": round(ach / SMEM_LOOKUP_PEAK, 4),
                               "how": f"{LOOKUPS_PER_BLOCK[wl]} lookups x blocks served by the table-driven warps (work-queue counters "
                                      f"of the last launch) / kernel time / {mhz:.0f} MHz / {N_SM} SMs; peak = measured conflict-free "
                                      "LDS.32 rate (profiles/r1_microbench.log); the bitsliced warps' share needs no lookups"}
    except Exception as e:
        lookup_roof = {"bound": "smem_lookup", "error": str(e)}

    # ---- the other BASELINE configs in the same driver record
    secondary = None
    if wl == "ctr128" and not args.no_secondary:
        secondary = {}
        ssteps, swarm = max(1, min(args.steps, 5)), 3
        for name, swl, sn in (("config2_ctr128_1GiB", "ctr128", min(nbytes, GIB)),
                              ("config3_xts256_512B_sectors", "xts256", nbytes),
                              ("config4_gcm128_4GiB_tag", "gcm128", min(nbytes, 4 * GIB))):
            try:
                m = measure(swl, sn, ssteps, swarm)
                ach = 2 * sn / (m["kernel_ms"] / 1e3) / 1e9
                secondary[name] = {"workload": f"{swl}, {sn / GIB:g} GiB per GPU" + (f", one message of {world * sn / GIB:g} GiB over {world} GPUs, 16-byte all-gather + combine in every step" if swl == "gcm128" and world > 1 else ""),
                                   "value": round(m["value"], 2), "unit": UNIT, "steps": ssteps, "warmup": swarm,
                                   "kernel_ms": round(m["kernel_ms"], 4), "achieved_GBps": round(ach, 1), "frac": round(ach / peak, 4),
                                   "gpu_launches": m["launches"], "parity_spot_check": m["parity"], **m["w"].extra,
                                   "kernel": KERNEL_NAMES[swl]}
            except Exception as e:
                secondary[name] = {"error": str(e)}
        # leave the headline state behind for the e2e leg
        uaes.fill_splitmix64(SEED, head["w"].first_block * 2, src, nbytes // 8)

    # ---- e2e: the reference-facing C ABI with HOST buffers, copies inside the timed region
    e2e = None
    try:
        if args.no_e2e or wl != "ctr128":
            raise RuntimeError("skipped (--no-e2e or secondary workload)")
        avail = mem_available_gib()
        e2e_gib = args.e2e_gib if args.e2e_gib else args.gib_per_gpu
        while e2e_gib > 0.25 and e2e_gib * world * 1.3 + 12 > avail:
            e2e_gib /= 2
        eb = int(e2e_gib * GIB)
        hbuf = torch.empty(eb + 16, dtype=torch.uint8, pin_memory=True)
        hbuf[:eb].copy_(src[:eb])
        torch.cuda.synchronize()
        shim = uaes.shim(128)
        hp = ctypes.c_void_p(hbuf.data_ptr())
        esteps = max(1, min(args.steps, args.e2e_steps))

        def e2e_time(fn, nb, steps=esteps):
            fn()                                               # warm-up (allocates staging chunks)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return round(world * nb * steps / GIB / float(tt.item()), 3)

        v = e2e_time(lambda: shim.AES_CTR_encrypt(key, iv, hp, eb, hp), eb)    # in place on the pinned host buffer
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": eb, "d2h_bytes_per_step": eb, "steps": esteps,
               "api": "AES_CTR_encrypt (libmicro_aes_128.so) on a pinned host buffer, in place",
               "buffer_gib_per_gpu": e2e_gib, "error": core.uaes_last_error()}
        # the same call as a real drop-in caller makes it: malloc'd (pageable) memory; and GCM (config 4)
        extras = {}
        try:
            pb = min(eb, 4 * GIB)
            page = np.empty(pb, dtype=np.uint8)
            page[:] = hbuf[:pb].numpy()
            pp = ctypes.c_void_p(page.ctypes.data)
            extras["ctr128_pageable"] = {"value": e2e_time(lambda: shim.AES_CTR_encrypt(key, iv, pp, pb, pp), pb, 2), "unit": UNIT,
                                         "buffer_gib_per_gpu": pb / GIB, "api": "AES_CTR_encrypt on a malloc'd (pageable) buffer, in place: "
                                         "pinned bounce chunks + helper threads inside the library"}
            del page
            gb = min(eb, 4 * GIB)
            extras["gcm128_pinned"] = {"value": e2e_time(lambda: shim.AES_GCM_encrypt(key, iv, None, 0, hp, gb, hp), gb, 2), "unit": UNIT,
                                       "buffer_gib_per_gpu": gb / GIB, "api": "AES_GCM_encrypt on a pinned host buffer, in place (one shard per "
                                       "staging chunk, contributions folded into the tag at the end)"}
            ndev = torch.cuda.device_count()
            if world == 1 and ndev > 1:                        # ONE process, ONE call, all GPUs of the box
                uaes.set_devices(0)
                extras["ctr128_pinned_all_devices"] = {"value": e2e_time(lambda: shim.AES_CTR_encrypt(key, iv, hp, eb, hp), eb), "unit": UNIT,
                                                       "devices": ndev, "api": "one AES_CTR_encrypt call spread over all GPUs (uaes_set_devices)"}
                uaes.set_devices(1)
            extras["error"] = core.uaes_last_error()
        except Exception as e:
            extras["error"] = str(e)
        e2e["extras"] = extras
        del hbuf
    except Exception as e:
        e2e = {"value": None, "unit": UNIT, "error": str(e)}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    kernel_ms = head["kernel_ms"]
    achieved = 2 * nbytes / (kernel_ms / 1e3) / 1e9        # 32 algorithmic bytes per 16-byte block
    traffic, traffic_src = None, None
    try:
        if wl == "ctr128" and nbytes == 16 * GIB:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ctr_traffic.json")))
            traffic, traffic_src = tj["dram_bytes_per_launch_16GiB"], "static: one ncu --set full capture, " + tj["source"]
    except Exception:
        pass
    cpu = cpu_throughput(slice_mib=args.cpu_slice_mib, reps=1) if not args.no_cpu and wl == "ctr128" else None
    cfg_name = {1: "the 16 GiB buffer BASELINE.json's metric names, on 1 B200"}.get(world, f"BASELINE config 5 shape: {world} x {args.gib_per_gpu:g} GiB sharded by counter range")
    line = {
        "metric": METRIC if wl == "ctr128" else f"{wl} throughput, {args.gib_per_gpu:g} GiB per B200", "value": round(head["value"], 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(head["total_ms"] / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"{'AES-128-CTR' if wl == 'ctr128' else wl}, {args.gib_per_gpu:g} GiB per GPU, out of place, device resident ({cfg_name})",
                   "bytes_per_gpu": nbytes, "total_bytes": nbytes * world,
                   "sharding": "contiguous keystream-block range per rank; NCCL broadcast of key||iv only" + ("; GCM: + one 16-byte all-gather per step" if wl == "gcm128" and world > 1 else ""),
                   "l2": f"inputs ({args.gib_per_gpu:g} GiB) larger than L2 (126 MB); no flush needed",
                   "input": f"splitmix64(seed=0x{SEED:x}) generated on device"},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": KERNEL_NAMES.get(wl, wl), "kernel_ms": round(kernel_ms, 4),
                     "algorithmic_bytes_per_launch": 2 * nbytes},
        "roofline_binding": lookup_roof, "queue_split": queue_split,
        "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": head["launches"],
        "parity_spot_check": head["parity"], "parity_ranks": world, **head["w"].extra,
        "secondary": secondary,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gib-per-gpu", type=float, default=16.0)
    ap.add_argument("--e2e-gib", type=float, default=0.0, help="host buffer for the e2e leg (default: auto)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-slice-mib", type=int, default=256)
    ap.add_argument("--workload", default="ctr128", choices=["ctr128", "ctr256", "ecb128", "ecb128dec", "xts256", "xts256dec", "gcm128", "gcmsiv128", "cbc128dec", "cfb128dec", "ocb128", "ccm128batch", "eax128batch", "siv128batch", "gcm128batch", "xts256unit"],
                    help="ctr128 is the headline (BASELINE.json metric); the others are the secondary configs")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip BASELINE configs 2-4 after the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

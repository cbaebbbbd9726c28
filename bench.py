#!/usr/bin/env python3
"""bench.py — BASELINE.json's headline metric on its quoted configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--streams S]

Workload (config C2): S = 4096 independent zlib streams per GPU, each one dynamic-Huffman
block inflating to exactly 65536 bytes of seeded word-soup text (workloads.c2_streams).
One "step" = one pass of batched inflate over all S streams.

  value     uncompressed GB/s, whole job (all ranks), inputs and outputs resident in HBM,
            CUDA events on the launching stream, max over ranks.
  e2e       same metric through the host-facing C-ABI call dcmp_inflate_batch on pinned HOST
            buffers: H2D of the compressed slab + offset tables and D2H of the inflated slab
            + result arrays are inside the timed region.
  roofline  algorithmic bytes (compressed read once + inflated written once) / kernel time
            vs the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU restatement of De.Inf.Ns.inflate (oracle/, "port") on this box's host
            cores, same streams.

`--impl reference` times that CPU path alone (the OCaml reference cannot be built in this
image: no ocaml/dune) and prints the same JSON shape with "impl": "reference".
Under torchrun rank 0 alone runs it.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402

STREAM_BYTES = 65536


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(streams):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
    `ncu --set full` summary (profiles/); only valid for the workload size it was captured on."""
    p = os.path.join(ROOT, "profiles", "inflate_parallel_ncu_summary.json")
    try:
        d = json.load(open(p))
        if int(d.get("streams", -1)) == int(streams):
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        pass
    return None


def make_workload(n, seed0):
    """Generate (or load from a /tmp cache) n C2 streams; returns packed numpy slabs + plaintext digest."""
    cache = os.path.join(tempfile.gettempdir(), "dcmp_c2_%d_%d.npz" % (n, seed0))
    if os.path.exists(cache):
        z = np.load(cache)
        return z["src"], z["so"], z["sl"], z["do"], z["dc"], z["adler"]
    from decompress_b200 import batch
    import zlib

    workers = max(1, min(32, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    comps, plains = workloads.c2_streams(n, STREAM_BYTES, seed0=seed0, workers=workers)
    src, so, sl, _dst, do, dc = batch.pack(comps, [STREAM_BYTES] * n, align=16)
    adler = np.array([zlib.adler32(p) for p in plains], dtype=np.uint32)
    try:
        np.savez(cache + ".tmp.npz", src=src, so=so, sl=sl, do=do, dc=dc, adler=adler)
        os.replace(cache + ".tmp.npz", cache)
    except OSError:
        pass
    return src, so, sl, do, dc, adler


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a
    thread (the timed region is tens of milliseconds, shorter than one `nvidia-smi -lms` period);
    falls back to the recipe's nvidia-smi line when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc, self.stop_flag, self.thread = gpu, [], None, False, None
        self.sm, self.reasons, self.max_sm, self.power = [], set(), None, []

    def _nvml_loop(self):
        import pynvml as N
        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
        h = N.nvmlDeviceGetHandleByIndex(idx)
        self.max_sm = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        self.ready = True
        while not self.stop_flag:
            if self.armed:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                r = get_reasons(h)
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
                try:
                    self.power.append(N.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        self.ready, self.armed = False, False
        try:
            import pynvml  # noqa: F401
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.ready and time.time() - t0 < 5 and self.thread.is_alive():
                time.sleep(0.005)
            if self.ready:
                return
        except Exception:
            pass
        self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def arm(self):
        """Called right before the timed region starts."""
        self.armed = True

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        self.armed = False
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "reasons": sorted(self.reasons), "samples": len(self.sm),
                    "power_w_max": round(max(self.power), 1) if self.power else None, "source": "nvml, 2 ms poll"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}


def cpu_baseline(src, so, sl, do, dc, budget_s=12.0):
    """Time the CPU restatement (oracle port) over the same streams with all host threads."""
    from tests import oracle

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    threads = os.cpu_count() or 1
    n = len(so)
    dst = np.zeros(int(do[-1] + dc[-1]) + 64, dtype=np.uint8)
    secs, st, prod = oracle.bench_batch(1, src, so, sl, dst, do, dc, threads)  # warm-up pass
    reps, total, t0 = 0, 0.0, time.time()
    times = []
    while reps < 5 and (time.time() - t0) < budget_s:
        secs, st, prod = oracle.bench_batch(1, src, so, sl, dst, do, dc, threads)
        times.append(secs)
        reps += 1
    assert int(st.max()) == 0
    med = float(np.median(times))
    return {"value": round(float(prod.sum()) / med / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": "%d of the same C2 streams (%d MiB inflated) per pass, median of %d passes, one stream per "
                      "thread over %d threads; CPU restatement of De.Inf.Ns.inflate (oracle/inflate_ns.c)"
                      % (n, int(prod.sum()) >> 20, reps, threads)}, med, float(prod.sum())


def run_reference(args, rank):
    """The reference arm: the reference algorithm's CPU implementation on the host cores."""
    if rank != 0:
        return
    n = args.streams
    src, so, sl, do, dc, _ = make_workload(n, 0)
    from tests import oracle

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    threads = os.cpu_count() or 1
    dst = np.zeros(int(do[-1] + dc[-1]) + 64, dtype=np.uint8)
    for _ in range(args.warmup):
        oracle.bench_batch(1, src, so, sl, dst, do, dc, threads)
    t = 0.0
    for _ in range(args.steps):
        secs, st, prod = oracle.bench_batch(1, src, so, sl, dst, do, dc, threads)
        assert int(st.max()) == 0
        t += secs
    val = float(prod.sum()) * args.steps / t / 1e9
    base = {"value": round(val, 4), "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": "all %d C2 streams (%d MiB inflated) per step, one stream per thread over %d threads; CPU "
                      "restatement of De.Inf.Ns.inflate (oracle/inflate_ns.c)" % (n, int(prod.sum()) >> 20, threads)}
    line = {"impl": "reference", "metric": "inflate_uncompressed_GBps", "value": round(val, 4), "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: %d zlib streams x 64 KiB, single dynamic-Huffman block each, CPU port of "
                                   "De.Inf.Ns.inflate on host cores (OCaml reference not buildable here)" % n,
                       "streams": n, "stream_bytes": STREAM_BYTES},
            "cpu_baseline": base,
            "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)



def run_c4(args, rank, local_rank, world):
    """BASELINE config C4 — Gz.Inf of concatenated gzip members (16..256 KiB payloads), STRONG scaling over the GPUs of
    one node: the whole file lives on rank 0's GPU; per step the member index is broadcast, every rank receives the
    contiguous slab of members `shard.partition` gives it (NCCL point-to-point over NVLink), inflates it with CRC-32 /
    ISIZE checked on the device, and its output slab goes back to rank 0 — in pieces, each sent while the next one is
    being decoded.  value = inflated bytes of the whole file / step time (max over ranks)."""
    import torch
    import torch.distributed as dist

    from decompress_b200 import _lib, batch, shard

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = _lib.Engine(local_rank)
    total = args.c4_mib << 20
    pieces_per_rank = max(1, args.c4_pieces)
    # ---- rank 0: the file and its index, resident in HBM before the timed region
    if rank == 0:
        t0 = time.time()
        cache = os.path.join(tempfile.gettempdir(), "dcmp_c4_%d" % args.c4_mib)
        if os.path.exists(cache + ".idx.npy") and os.path.exists(cache + ".blob.npy"):
            index, blob_h = np.load(cache + ".idx.npy"), np.load(cache + ".blob.npy")
            sizes = index[:, 2]
        else:
            members, sizes = workloads.c4_blob(total, seed=0)
            lens = np.array([len(m) for m in members], dtype=np.int64)
            offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
            blob_h = np.frombuffer(b"".join(members), dtype=np.uint8)
            del members
            index = np.stack([offs, lens, sizes], axis=1)
            try:  # (the 1/2/4/8-GPU runs of one session share the file)
                np.save(cache + ".idx.npy", index)
                np.save(cache + ".blob.npy", blob_h)
            except OSError:
                pass
        blob = torch.from_numpy(blob_h.copy()).to(dev)
        out_full = torch.empty(int(sizes.sum()) + 64, dtype=torch.uint8, device=dev)
        gen_s = time.time() - t0
    else:
        index, blob, out_full, gen_s = None, None, None, 0.0
    # peer transport: rank 0's two buffers are mapped into every other process (CUDA IPC); slabs then move with plain
    # device-to-device copies — the copy engines over NVLink / NVSwitch, no SM spent on the transfer — and the only
    # collectives left are the index broadcast and one barrier that closes the step
    peer = world > 1 and args.c4_transport == "peer"
    if peer:
        # (the member index is part of the file: it sits in rank 0's memory too and is read through the same mapping)
        handles = [None, None, None, None]
        if rank == 0:
            index_dev = torch.from_numpy(np.ascontiguousarray(index, dtype=np.int64)).to(dev).view(torch.uint8).reshape(-1)
            handles = [blob.untyped_storage()._share_cuda_(), out_full.untyped_storage()._share_cuda_(),
                       index_dev.untyped_storage()._share_cuda_(), tuple(index.shape)]
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            def _open(h):
                st_ = torch.UntypedStorage._new_shared_cuda(*h)
                return torch.empty(0, dtype=torch.uint8, device=st_.device).set_(st_)
            blob, out_full, index_dev = _open(handles[0]), _open(handles[1]), _open(handles[2])
        idx_shape = handles[3]
        idx_bytes = int(idx_shape[0]) * int(idx_shape[1]) * 8
        idx_local = torch.empty(idx_bytes, dtype=torch.uint8, device=dev)
        idx_pinned = torch.empty(idx_bytes, dtype=torch.uint8).pin_memory()
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        token = torch.zeros(1, device=dev)

    def fetch_index():
        """The member index on this rank's host: NCCL broadcast, or — peer transport — a copy out of rank 0's mapped memory."""
        if world == 1:
            return index
        if not peer:
            return shard.broadcast_index(index, 0, dev)
        idx_local.copy_(index_dev[:idx_bytes], non_blocking=True)      # NVLink (rank 0: a local copy)
        idx_pinned.copy_(idx_local, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return idx_pinned.numpy().view(np.int64).reshape(idx_shape)

    def one_step(timers=None):
        ev = lambda: torch.cuda.Event(enable_timing=True)
        marks = [ev() for _ in range(4)]
        marks[0].record()
        if peer:
            # Plan on the device: the index comes over NVLink out of rank 0's mapping, prefix sums and the two levels of cuts
            # (ranks, then this rank's pieces) are a handful of small kernels, and only 3 x (world + 1) + 3 x (pieces + 1)
            # integers ever reach the host (round 2, first version: the whole 2.3 MB index went to the host and numpy
            # planned there, 5.1 ms of a 17.4 ms step at 8 GPUs).
            idx_local.copy_(index_dev[:idx_bytes], non_blocking=True)
            I = idx_local.view(torch.int64).view(int(idx_shape[0]), int(idx_shape[1]))
            n = I.shape[0]
            z1 = torch.zeros(1, dtype=torch.int64, device=dev)
            ends_c_d = torch.cat([z1, torch.cumsum(I[:, 1], 0)])
            ends_o_d = torch.cat([z1, torch.cumsum(I[:, 2], 0)])
            cuts_d = shard.partition_cuts(I[:, 1] + I[:, 2], world)
            head = torch.stack([cuts_d, ends_c_d[cuts_d], ends_o_d[cuts_d]]).cpu().numpy()
            parts = [(int(head[0][r]), int(head[0][r + 1])) for r in range(world)]
            ranges_c = [(int(head[1][r]), int(head[1][r + 1])) for r in range(world)]
            ranges_o = [(int(head[2][r]), int(head[2][r + 1])) for r in range(world)]
            job_src, job_out = int(head[1][world]), int(head[2][world])
            a, b = parts[rank]
            m = b - a
            my_cuts, pc_c, pc_o = [], {}, {}
            if m > 0:
                pc_d = shard.partition_cuts(I[a:b, 2], min(pieces_per_rank, m))
                pinfo = torch.stack([pc_d, ends_c_d[a + pc_d], ends_o_d[a + pc_d]]).cpu().numpy()
                for i in range(pinfo.shape[1] - 1):
                    x, y = int(pinfo[0][i]), int(pinfo[0][i + 1])
                    if y > x:
                        my_cuts.append((x, y))
                        pc_c[x], pc_c[y], pc_o[x], pc_o[y] = int(pinfo[1][i]), int(pinfo[1][i + 1]), int(pinfo[2][i]), int(pinfo[2][i + 1])
            npieces = len(my_cuts)
            scatter_works = [[] for _ in range(npieces)]
            in_events = [None] * npieces
            if rank == 0:
                local = blob[ranges_c[0][0]:ranges_c[0][1]]
            else:
                local = torch.empty(ranges_c[rank][1] - ranges_c[rank][0] + 64, dtype=torch.uint8, device=dev)
                s_in.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_in):
                    for p in range(npieces):            # pull this rank's compressed slab out of rank 0's memory, piece by piece
                        x, y = my_cuts[p]
                        local[pc_c[x] - ranges_c[rank][0]:pc_c[y] - ranges_c[rank][0]].copy_(blob[pc_c[x]:pc_c[y]], non_blocking=True)
                        in_events[p] = torch.cuda.Event()
                        in_events[p].record(s_in)
            marks[1].record()
            so = (I[a:b, 0] - ends_c_d[a]).contiguous()
            sl = I[a:b, 1].contiguous()
            do = (ends_o_d[a:b] - ends_o_d[a]).contiguous()
            dc = I[a:b, 2].contiguous()
        else:
          idx = fetch_index()
          n = idx.shape[0]
          parts = shard.partition(idx[:, 1] + idx[:, 2], world)
          ends_c = np.concatenate([[0], np.cumsum(idx[:, 1])])
          ends_o = np.concatenate([[0], np.cumsum(idx[:, 2])])
          ranges_c = [(int(ends_c[a]), int(ends_c[b])) for a, b in parts]
          ranges_o = [(int(ends_o[a]), int(ends_o[b])) for a, b in parts]
          a, b = parts[rank]
          m = b - a
          # pieces: contiguous member ranges of about equal output.  Compressed slabs arrive piece by piece and a piece is
          # decoded as soon as it is there; its output leaves for rank 0 while the next piece is being decoded.
          def cuts_of(r):
              ra, rb = parts[r]
              if rb == ra:
                  return []
              ps = shard.partition(idx[ra:rb, 2], min(pieces_per_rank, rb - ra))
              return [(x, y) for x, y in ps if y > x]
          all_cuts = [cuts_of(r) for r in range(world)]
          my_cuts = all_cuts[rank]
          npieces = max(len(c) for c in all_cuts) if all_cuts else 0
          scatter_works = [[] for _ in range(npieces)]
          in_events = [None] * npieces
          if world > 1:
              if rank == 0:
                  local = blob[ranges_c[0][0]:ranges_c[0][1]]
              else:
                  local = torch.empty(ranges_c[rank][1] - ranges_c[rank][0] + 64, dtype=torch.uint8, device=dev)
              for p in range(npieces):
                  ops = []
                  if rank == 0:
                      for r in range(1, world):
                          if p < len(all_cuts[r]):
                              x, y = all_cuts[r][p]
                              ra = parts[r][0]
                              ops.append(dist.P2POp(dist.isend, blob[int(ends_c[ra + x]):int(ends_c[ra + y])], r))
                  elif p < len(my_cuts):
                      x, y = my_cuts[p]
                      ops.append(dist.P2POp(dist.irecv, local[int(ends_c[a + x] - ends_c[a]):int(ends_c[a + y] - ends_c[a])], 0))
                  if ops:
                      scatter_works[p] = dist.batch_isend_irecv(ops)
          else:
              local = blob[ranges_c[0][0]:ranges_c[0][1]]
          marks[1].record()
          do_h = (ends_o[a:b] - ends_o[a]).astype(np.int64)
          tabs = torch.from_numpy(np.stack([idx[a:b, 0] - ends_c[a], idx[a:b, 1], do_h, idx[a:b, 2]]).astype(np.int64)).to(dev)  # one upload
          so, sl, do, dc = tabs[0], tabs[1], tabs[2], tabs[3]
          job_src, job_out = int(ends_c[-1]), int(ends_o[-1])
        my_out = out_full[ranges_o[0][0]:ranges_o[0][1]] if rank == 0 else torch.empty(ranges_o[rank][1] - ranges_o[rank][0] + 64, dtype=torch.uint8, device=dev)
        cons = torch.zeros(max(m, 1), dtype=torch.int64, device=dev)
        prod = torch.zeros(max(m, 1), dtype=torch.int64, device=dev)
        st = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        works = []
        for p in range(npieces):
            if p < len(my_cuts):
                if rank != 0:
                    if peer:
                        torch.cuda.current_stream().wait_event(in_events[p])
                    for w in scatter_works[p]:
                        w.wait()           # (stream-ordered: the host does not block)
                x, y = my_cuts[p]
                batch.inflate_device(_lib.FMT_GZIP, local, so[x:y], sl[x:y], my_out, do[x:y], dc[x:y], cons[x:y], prod[x:y], st[x:y], engine=eng)
                if peer and rank != 0:                  # the decoded piece goes straight into its place in rank 0's memory
                    evd = torch.cuda.Event()
                    evd.record()
                    s_out.wait_event(evd)
                    lo, hi = pc_o[x] - ranges_o[rank][0], pc_o[y] - ranges_o[rank][0]
                    with torch.cuda.stream(s_out):
                        out_full[ranges_o[rank][0] + lo:ranges_o[rank][0] + hi].copy_(my_out[lo:hi], non_blocking=True)
            if world > 1 and not peer:
                ops = []
                if rank == 0:
                    for r in range(1, world):
                        if p < len(all_cuts[r]):
                            x, y = all_cuts[r][p]
                            ra = parts[r][0]
                            ops.append(dist.P2POp(dist.irecv, out_full[int(ends_o[ra + x]):int(ends_o[ra + y])], r))
                elif p < len(my_cuts):
                    x, y = my_cuts[p]
                    ops.append(dist.P2POp(dist.isend, my_out[int(do_h[x]):int(do_h[y - 1] + idx[a + y - 1, 2])], 0))
                if ops:
                    works += dist.batch_isend_irecv(ops)
        if rank == 0:
            for ws in scatter_works:
                works += ws
        if peer:  # the step is over when every rank's copies have landed: one collective, ordered after them on each rank
            torch.cuda.current_stream().wait_stream(s_out)
            torch.cuda.current_stream().wait_stream(s_in)
        marks[2].record()
        for w in works:
            w.wait()
        if peer:
            dist.all_reduce(token)
        marks[3].record()
        if timers is not None:
            torch.cuda.synchronize()
            timers.append([marks[k].elapsed_time(marks[k + 1]) for k in range(3)])
        return st, prod, m, n, job_src, job_out

    W = max(3, args.warmup)
    for _ in range(W):
        st, prod, m, n, job_src, job_out = one_step()
    torch.cuda.synchronize()
    assert m == 0 or int(st[:m].abs().sum().item()) == 0, "a gzip member failed (CRC-32 / ISIZE are checked on the device)"
    s0 = eng.stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.arm()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        st, prod, m, n, job_src, job_out = one_step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    s1 = eng.stats()
    assert m == 0 or (int(st[:m].abs().sum().item()) == 0 and int(prod[:m].sum().item()) > 0)
    assert s1["inflate_redone"] == s0["inflate_redone"], "a valid member fell back to the general kernel"
    phases = []
    for _ in range(3):
        one_step(phases)     # untimed: where a step's time goes (scatter | decode (+ overlapped sends) | waiting for the last pieces)
    ph = np.median(np.array(phases), axis=0)
    if rank == 0:  # every byte of the gathered file is checked once, after the timed region: CRC-32 of each member's payload
        import zlib
        outh = out_full[:job_out].cpu().numpy()
        e_o = np.concatenate([[0], np.cumsum(index[:, 2])])
        step = max(1, n // 512)
        for k in list(range(0, n, step)) + [n - 1]:
            tr = blob_h[index[k, 0] + index[k, 1] - 8: index[k, 0] + index[k, 1] - 4]
            assert zlib.crc32(outh[e_o[k]:e_o[k + 1]].tobytes()) == int.from_bytes(tr.tobytes(), "little"), "member %d" % k
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
        pp = torch.tensor(ph, dtype=torch.float64, device=dev)
        dist.all_reduce(pp, op=dist.ReduceOp.MAX)
        ph = pp.cpu().numpy()
    if rank == 0:
        peak, peak_src = load_peaks()
        ms_step = ms / args.steps
        value = job_out / (ms_step * 1e-3) / 1e9
        line = {"metric": "inflate_uncompressed_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": "C4: Gz.Inf of %d MiB of concatenated gzip members (%d members, 16..256 KiB payloads of "
                                       "synthetic text, ratio %.2f:1), the file resident on rank 0's GPU, sharded over %d GPU(s) by member "
                                       "ranges" % (args.c4_mib, n, job_out / job_src, world),
                           "members": int(n), "compressed_bytes": int(job_src), "inflated_bytes": int(job_out),
                           "transport": args.c4_transport if world > 1 else "none",
                           "parallelism": "index broadcast + scatter of compressed slabs + gather of output slabs (peer: copy engines over NVLink through "
                                          "CUDA IPC mappings of rank 0's buffers; nccl: NCCL send/recv), "
                                          "both in %d pieces per rank: a piece is decoded as soon as it has arrived and sent back while "
                                          "the next one is decoded" % pieces_per_rank,
                           "l2": "working set %d MiB per step > 126 MB L2" % ((job_src + job_out) >> 20)},
                "roofline": {"bound": "hbm", "achieved": round((job_src + job_out) / world / (ph[1] * 1e-3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                             "frac": round((job_src + job_out) / world / (ph[1] * 1e-3) / 1e9 / peak, 4), "traffic": None,
                             "kernel": "gzip_header + inflate_parallel + gzip_trailer over one rank's share, timed as the decode phase "
                                       "of a step (median of 3 untimed steps)", "peak_source": peak_src},
                "phases_ms": {"index_broadcast_and_plan": round(float(ph[0]), 3), "decode_with_overlapped_transfers": round(float(ph[1]), 3),
                              "waiting_for_the_last_pieces": round(float(ph[2]), 3)},
                "collective_share": round(float(ph[0] + ph[2]) / float(ph.sum()), 3),
                "gpu_launches": int(launches), "clocks": clocks, "generated_in_s": round(gen_s, 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU (C2 = 4096)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"], help="c2: BASELINE's quoted config (default); c4: multi-member gzip, strong scaling")
    ap.add_argument("--c4-mib", type=int, default=8192, help="inflated size of the C4 file (BASELINE: 8 GiB)")
    ap.add_argument("--c4-pieces", type=int, default=4, help="output pieces per rank (a piece is sent while the next is decoded)")
    ap.add_argument("--c4-transport", default="peer", choices=["peer", "nccl"],
                    help="how slabs travel: peer = copy engines over NVLink through CUDA IPC mappings of rank 0's buffers; nccl = NCCL send/recv")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "c4":
        run_c4(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist

    from decompress_b200 import _lib, batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libdcmp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    n = args.streams
    src, so, sl, do, dc, adler = make_workload(n, rank * n)
    total_src = int(sl.sum())
    total_out = n * STREAM_BYTES
    eng = _lib.Engine(local_rank)

    # ---- device-resident arm -----------------------------------------------------------------
    t = lambda a: torch.from_numpy(a).to(dev)
    d_src, d_so, d_sl, d_do, d_dc = t(src), t(so), t(sl), t(do), t(dc)
    d_dst = torch.zeros(int(do[-1] + dc[-1]) + 64, dtype=torch.uint8, device=dev)
    d_cons = torch.zeros(n, dtype=torch.int64, device=dev)
    d_prod = torch.zeros(n, dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)

    def step_device():
        batch.inflate_device(batch.FMT_ZLIB, d_src, d_so, d_sl, d_dst, d_do, d_dc, d_cons, d_prod, d_st, engine=eng)

    for _ in range(W):
        step_device()
    torch.cuda.synchronize()
    assert int(d_st.abs().sum().item()) == 0 and int(d_prod.sum().item()) == total_out, "inflate failed in warm-up"
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.arm()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    # result check after the timed region: Adler-32 of every inflated stream equals the plaintext's
    out = d_dst.cpu().numpy()
    import zlib
    for k in range(0, n, max(1, n // 64)):
        assert zlib.adler32(out[do[k]:do[k] + STREAM_BYTES].tobytes()) == int(adler[k]), "wrong output for stream %d" % k
    assert int(d_st.abs().sum().item()) == 0

    # ---- end-to-end arm: host pinned buffers through dcmp_inflate_batch ------------------------
    h_src = torch.from_numpy(src).pin_memory()
    h_dst = torch.zeros(int(do[-1] + dc[-1]) + 64, dtype=torch.uint8).pin_memory()
    cons = np.zeros(n, dtype=np.int64)
    prod = np.zeros(n, dtype=np.int64)
    stat = np.zeros(n, dtype=np.int32)
    P = lambda a: C.c_void_p(a.ctypes.data)

    def step_e2e():
        rc = _lib.lib().dcmp_inflate_batch(eng.handle, batch.FMT_ZLIB, n, C.c_void_p(h_src.data_ptr()), P(so), P(sl),
                                           C.c_void_p(h_dst.data_ptr()), P(do), P(dc), P(cons), P(prod), P(stat))
        _lib.check(rc, "dcmp_inflate_batch")

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert int(np.abs(stat).sum()) == 0 and int(prod.sum()) == total_out
    hv = h_dst.numpy()
    assert zlib.adler32(hv[do[n - 1]:do[n - 1] + STREAM_BYTES].tobytes()) == int(adler[n - 1])

    # ---- what the box allows: the same bytes as plain pinned copies, both directions at once, all ranks at the same time.
    # e2e cannot beat this: it is the PCIe / host-memory ceiling of THIS box at THIS number of ranks.
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    d_tmp_in = torch.empty(h_src.numel(), dtype=torch.uint8, device=dev)

    def step_copy():
        with torch.cuda.stream(s_h2d):
            d_tmp_in.copy_(h_src, non_blocking=True)
        with torch.cuda.stream(s_d2h):
            h_dst[:total_out].copy_(d_dst[:total_out], non_blocking=True)
        s_h2d.synchronize()
        s_d2h.synchronize()

    step_copy()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_copy()
    copy_s = time.perf_counter() - t0
    # ---- the same call on PAGEABLE buffers (what an OCaml Bigarray is unless it is registered with dcmp_host_register)
    p_src, p_dst = src.copy(), np.zeros(int(do[-1] + dc[-1]) + 64, dtype=np.uint8)

    def step_pageable():
        rc = _lib.lib().dcmp_inflate_batch(eng.handle, batch.FMT_ZLIB, n, P(p_src), P(so), P(sl), P(p_dst), P(do), P(dc), P(cons), P(prod), P(stat))
        _lib.check(rc, "dcmp_inflate_batch")

    step_pageable()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        step_pageable()
    page_s = (time.perf_counter() - t0) / 3
    assert int(np.abs(stat).sum()) == 0 and zlib.adler32(p_dst[do[0]:do[0] + STREAM_BYTES].tobytes()) == int(adler[0])

    # ---- reduce over ranks (max time) ------------------------------------------------------------
    if world > 1:
        tt = torch.tensor([ms, e2e_s, copy_s, page_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s, copy_s, page_s = float(tt[0]), float(tt[1]), float(tt[2]), float(tt[3])
        tb = torch.tensor([float(total_src), float(total_out)], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        job_src, job_out = float(tb[0]), float(tb[1])
    else:
        job_src, job_out = float(total_src), float(total_out)

    if rank == 0:
        peak, peak_src = load_peaks()
        ms_step = ms / args.steps
        value = job_out / (ms_step * 1e-3) / 1e9
        # roofline is per GPU / per launch: this rank's algorithmic bytes over the average launch duration
        achieved = (total_src + total_out) / (ms_step * 1e-3) / 1e9
        line = {
            "metric": "inflate_uncompressed_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: %d independent zlib streams per GPU, each one dynamic-Huffman block "
                                   "inflating to 65536 B of seeded word-soup text (ratio %.2f:1)"
                                   % (n, total_out / total_src),
                       "streams_per_gpu": n, "stream_bytes": STREAM_BYTES, "compressed_bytes_per_gpu": total_src,
                       "parallelism": "streams sharded across GPUs, no data-path collective",
                       "l2": "working set %d MiB per step > 126 MB L2 (no flush needed)" % ((total_src + total_out) >> 20)},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": load_traffic(n),
                         "algorithmic_bytes_per_launch": int(total_src + total_out),
                         "kernel": "dcmp::inflate_parallel_kernel (compressed bytes read once + inflated bytes written "
                                   "once per launch; the general kernel that follows it finds no stream to redo)",
                         "peak_source": peak_src},
            "e2e": {"value": round(job_out / (e2e_s / e2e_steps) / 1e9, 3), "unit": "GB/s",
                    "h2d_bytes_per_step": int(total_src + 4 * 8 * n), "d2h_bytes_per_step": int(total_out + 20 * n),
                    "steps": e2e_steps, "api": "dcmp_inflate_batch (pinned host buffers)",
                    "pcie_ceiling": round(job_out / (copy_s / e2e_steps) / 1e9, 3),
                    "pcie_ceiling_how": "the step's H2D and D2H bytes as plain pinned cudaMemcpyAsync on two streams, every rank at the same "
                                        "time: the same metric with no kernel at all",
                    "pageable": round(job_out / page_s / 1e9, 3)},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            base, _, _ = cpu_baseline(src, so, sl, do, dc)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

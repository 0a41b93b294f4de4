#!/usr/bin/env python
"""bench.py -- active-voxels/s through the SG-NN generator (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                 B200 arm (this repo's engine)
  python bench.py --impl reference --gpus N --steps K --warmup W reference arm: the CPU path
                                                                 (oracle port of model.py + restated scn)

One step = one pass of the hot path (GenModel forward, loss_weights = ones(5), eval) over one batch of
`--blocks` synthetic 64^3 TSDF blocks @5 % per GPU (BASELINE.json configs[1]; SURVEY §8(d) inputs).
Weak scaling: every rank owns its own blocks (block i -> rank i mod N), one NCCL broadcast of the weights,
no data-path collective.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'input_active_voxels_per_s_through_sgnn_generator'
UNIT = 'voxels/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--blocks', type=int, default=0, help='64^3 blocks per GPU per step (0: 32, or 512 for --config 3)')
    ap.add_argument('--sets', type=int, default=4, help='distinct input batches rotated across steps')
    ap.add_argument('--param-seed', type=int, default=0)
    ap.add_argument('--cpu-blocks', type=int, default=0, help='blocks per CPU forward: 0 = 8 for the cpu_baseline sample of the '
                    'B200 arm, --blocks (the same workload) for --impl reference')
    ap.add_argument('--cpu-reps', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ledger', default='', help='write the per-layer ledger JSON here')
    ap.add_argument('--config', type=int, default=1, choices=[1, 2, 3, 4],
                    help='BASELINE.json configs[i]: 1 = 32 blocks per GPU (default), 3 = 512 blocks per GPU (4096 over 8 GPUs), '
                         '2 = one 128^3 block bf16 Convolution+Deconvolution, 4 = 10 M-site rulebook build')
    ap.add_argument('--conv-mode', default='tc32', choices=['exact', 'tc32'],
                    help="exact: fixed-order FFMA convolutions; tc32: Cout=16 convolutions on tcgen05 (3-way bf16 split)")
    ap.add_argument('--tc32-min-rows', type=int, default=0, help='GenModel.tc32_min_rows (A/B runs; 0 = default)')
    ap.add_argument('--overlap-max-rows', type=int, default=0, help='A/B: GenModel.overlap_max_rows (0 default, -1 never)')
    ap.add_argument('--dense-rules', action='store_true', help='A/B: dense neighbour table on the encoder input level')
    ap.add_argument('--ur-min-rows', type=int, default=0, help='GenModel.ur_min_rows (A/B runs; 0 = default)')
    return ap.parse_args()


def load_synth():
    """sgnn_b200/synth.py by FILE PATH: importing it as part of the package would run sgnn_b200/__init__.py and map the product
    libsgnn_b200.so into the reference arm's process (numpy / torch only, no package-relative imports)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('sgnn_synth', os.path.join(ROOT, 'sgnn_b200', 'synth.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------ reference arm
def cpu_threads():
    """Threads for the CPU arm: all host cores up to 16 -- beyond that the restated scn path (many small torch ops)
    gets SLOWER from intra-op oversubscription (69 s per block with 128 threads on the GPU box vs 0.5 s with 8)."""
    return max(1, min(os.cpu_count() or 1, 16))


def cpu_forward_rate(n_blocks, reps, seed, first_block=0):
    """Times the oracle port of the reference CPU path (oracle/genmodel.py on oracle/sparseconvnet).
    This is the ONE place bench.py executes oracle code: as the thing the B200 arm is compared with."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from genmodel import OracleGenModel
    synth = load_synth()
    fill_parameters, synthetic_batch = synth.fill_parameters, synth.synthetic_batch
    cores = cpu_threads()
    torch.set_num_threads(cores)
    m = OracleGenModel()
    fill_parameters(m, seed)
    m.eval()
    locs, feats = synthetic_batch(n_blocks, 64, 0.05, first=first_block)
    times = []
    t_start = time.perf_counter()
    with torch.no_grad():
        m(locs, feats)                      # warm-up (allocator, thread pool)
        for _ in range(reps):
            t0 = time.perf_counter()
            m(locs, feats)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > 45:      # bounded sample: stop after ~45 s of CPU work
                break
    t = float(np.median(times))
    return locs.shape[0] / t, t, cores, locs.shape[0]


def run_reference(args):
    """Reference arm: the CPU path (oracle port of model.py on the restated SparseConvNet-CPU ops) on the SAME workload as the
    B200 arm -- `--blocks` 64^3 blocks per step, exactly --steps timed steps after --warmup warm-ups (3
    warm-ups, median of >= 10 also reported), all usable host threads.  A 9-minute budget bounds a slow box: steps not run
    are reported as such.  This process never loads the product library."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from genmodel import OracleGenModel
    synth = load_synth()
    cores = cpu_threads()
    torch.set_num_threads(cores)
    m = OracleGenModel()
    synth.fill_parameters(m, args.param_seed)
    m.eval()
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    blocks = args.cpu_blocks if args.cpu_blocks > 0 else args.blocks
    batches = [synth.synthetic_batch(blocks, 64, 0.05, first=s * blocks) for s in range(args.sets)]
    times, vox = [], []
    t_begin = time.perf_counter()
    with torch.no_grad():
        for s in range(warm + steps):
            locs, feats = batches[s % args.sets]
            t0 = time.perf_counter()
            m(locs, feats)
            dt = time.perf_counter() - t0
            if s >= warm:
                times.append(dt)
                vox.append(locs.shape[0])
            if time.perf_counter() - t_begin > 540 and len(times) >= 3:
                break
    t_total = float(sum(times))
    value = sum(vox) / t_total
    done = len(times)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': done, 'warmup': warm, 'ms_per_step': 1e3 * t_total / done, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1]: %d synthetic 64^3 TSDF blocks @5%% per step, full 3-level coarse-to-fine '
                               'generator, fp32' % blocks, 'blocks_per_step': blocks, 'host': 'cpu',
                   'steps_requested': steps, 'median_ms_per_step': 1e3 * float(np.median(times)),
                   'product_library_loaded': any('libsgnn_b200' in l for l in open('/proc/self/maps'))},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d timed forward(s) of %d block(s) after %d warm-up(s); restated SparseConvNet-CPU path '
                                   '(oracle O2) under the oracle port of model.py' % (done, blocks, warm)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm
class ClockSampler(threading.Thread):
    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
                self.rows.append([v.strip() for v in out.stdout.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': float(rows[0][1]), 'reasons': reasons,
                'samples': len(rows)}


class ConvProfiler(object):
    """Event pair around every sgnn_conv_forward of the timed steps + the algorithmic-byte ledger
    (bytes = (N_in*Cin + N_out*Cout)*4 + 8*R + K*Cin*Cout*4, flops = 2*R*Cin*Cout)."""

    def __init__(self):
        self.events, self.ledger, self.count_rules = [], [], False

    class _Ctx(object):
        def __init__(self, prof, rec):
            self.prof, self.rec = prof, rec
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

        def done(self):
            self.e1.record()
            self.prof.events.append((self.e0, self.e1, self.rec))

    def conv(self, x, nbr, weight, n_out, child_mode, has_res, has_b):
        K, cin, cout = weight.shape
        rec = {'op': 'conv_child' if child_mode else ('conv_k27' if K == 27 else 'conv_k8'),
               'n_in': int(x.shape[0]), 'n_out': n_out, 'cin': cin, 'cout': cout, 'K': K}
        if self.count_rules:
            rec['R'] = child_rule_count(nbr) if child_mode else int((nbr >= 0).sum().item())
            self.ledger.append(rec)
        return ConvProfiler._Ctx(self, rec)


def kernel_source_hash():
    """sha256 over the CUDA sources: profiles/conv_dram_traffic.json (an ncu capture) is only quoted for the kernels it measured."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, 'sgnn_b200', 'csrc', '*.cu*'))):
        h.update(open(f, 'rb').read())
    return h.hexdigest()


def child_rule_count(nbr):
    """Rules of the child-mode convolution = for each parent/child/offset with an existing parent neighbour."""
    present = (nbr >= 0).view(3, 3, 3, -1).float()           # [pz,py,px, n]
    # per axis a child at c in {0,1} reaches parent offsets {-1,0,0} (c=0) or {0,0,1} (c=1): weights (1,2,0),(0,2,1)
    wz = torch.tensor([[1., 2., 0.], [0., 2., 1.]], device=nbr.device)
    tot = torch.einsum('az,by,cx,zyxn->', wz, wz, wz, present)
    return int(tot.item())


def run_b200(args):
    import torch.distributed as dist
    import sgnn_b200
    import sgnn_b200.engine as E
    from sgnn_b200 import shard
    from sgnn_b200._lib import lib
    from sgnn_b200.synth import fill_parameters, synthetic_batch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: the B200 arm needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # one slice of the host cores per rank: the host side of the e2e path (pinned copies, launches) otherwise migrates
    # between cores shared by all ranks (round 1: e2e scaling 0.917 at N = 8 with every GPU reporting the same CPU affinity)
    bound = None
    try:
        cores_all = sorted(os.sched_getaffinity(0))
        lw = int(os.environ.get('LOCAL_WORLD_SIZE', str(world)))
        if lw > 1 and len(cores_all) >= 2 * lw:
            per = len(cores_all) // lw
            mine = cores_all[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            torch.set_num_threads(max(1, min(per, 8)))
            bound = '%d-%d' % (mine[0], mine[-1])
    except Exception:
        pass
    if world > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'      # keep NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group('nccl', device_id=dev)
    ones = np.ones(5, dtype=np.float32)

    model = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
    if rank == 0:
        fill_parameters(model, args.param_seed)
    model = model.to(dev).eval()
    shard.broadcast_parameters(model, src=0)          # the one collective of the path (2.57 MB)
    model.return_long = True
    model.conv_mode = args.conv_mode
    model.tc32_min_rows, model.ur_min_rows = args.tc32_min_rows, args.ur_min_rows
    model.dense_rules = bool(args.dense_rules)
    model.overlap_max_rows = int(args.overlap_max_rows)

    if args.blocks <= 0:
        args.blocks = 512 if args.config == 3 else 32
    # inputs: `sets` distinct batches per rank (global block id = (set*world + rank)*blocks + i), resident in HBM
    host, resident = [], []
    for s in range(args.sets):
        locs, feats = synthetic_batch(args.blocks, 64, 0.05, first=(s * world + rank) * args.blocks)
        host.append((locs.to(torch.int16).pin_memory(), feats.pin_memory()))   # coordinates cross PCIe as int16 x 4
        resident.append((locs.to(dev), feats.to(dev)))
    vox = [int(h[0].shape[0]) for h in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(i):
        locs, feats = resident[i % args.sets]
        return model([locs, feats, args.blocks], ones)      # scn's [coords, features, batch_size] input form

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds the byte ledger of set 0 on rank 0)
    prof = ConvProfiler()
    for w in range(max(args.warmup, 3)):
        step(w)
    torch.cuda.synchronize()
    ledger = []
    if rank == 0:
        prof.count_rules = True
        E.PROFILER = prof
        out = model.forward_fused(list(resident[0]), ones)     # same arithmetic, every conv visible to the hook
        torch.cuda.synchronize()
        E.PROFILER = None
        prof.count_rules = False
        ledger = prof.ledger
        prof.events, prof.ledger = [], []
        final_rows = int(out[0][0].shape[0])
        level_rows = [int(l[0].shape[0]) if not isinstance(l[0], list) else 0 for l in out[1]]
    # ---- timed region: exactly K steps, L2 flushed before each (flush inside the bracket, ~40 us of 256 MiB write)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = lib.sgnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        step(i)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    launches = (lib.sgnn_launch_count() - l0) / max(args.steps, 1)
    dev_ms = e0.elapsed_time(e1)
    # ---- same K steps again with a CUDA-event pair around every convolution launch (native SGNN_GEN_PROFILE):
    # the dominant kernel's live duration.  Kept out of the bracket above because it adds one sync per step.
    conv_ms, n_conv, prof_ms = 0.0, 0, 0.0
    conv_table = []
    phase_table = []
    if rank == 0 and getattr(model, '_native', None) is not None:
        model._native.profile = True
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            step(i)
            conv_ms += model._native.last.conv_ms
            n_conv += model._native.last.n_conv
        p1.record()
        import ctypes as _C
        rec6, rms = (_C.c_int64 * 6)(), _C.c_float(0)
        for ci in range(int(model._native.last.n_conv)):     # per-convolution times of the last profiled step
            if lib.sgnn_generator_profile_entry(ci, rec6, _C.byref(rms)) != 0:
                break
            conv_table.append({'n_out': rec6[0], 'cin': rec6[1], 'cout': rec6[2], 'K': rec6[3], 'child': rec6[4],
                               'tc': rec6[5], 'us': round(rms.value * 1e3, 2)})
        torch.cuda.synchronize()
        prof_ms = p0.elapsed_time(p1)
        model._native.profile = False
        # ---- and one pass with a CUDA event at every phase boundary (SGNN_GEN_PHASES): where the step goes, in situ
        model._native.phases = True
        flush.fill_(1)
        step(0)
        torch.cuda.synchronize()
        phase_table = [[name, round(ms * 1e3, 1)] for name, ms in model._native.phase_table()]
        model._native.phases = False
    ms = shard.max_over_ranks(dev_ms, dev)
    vox_timed = sum(vox[i % args.sets] for i in range(args.steps))
    total_vox = shard.sum_over_ranks(vox_timed, dev)
    value = total_vox / (ms * 1e-3)

    # ---- e2e: HOST buffers in, host result out, through the public host-buffer API (sgnn_b200.streaming): per step the
    # pinned H2D of that step's coords + features and the D2H of its result coords + TSDF are inside the timed region
    # (they overlap the neighbouring steps' compute on the copy streams).
    from sgnn_b200.streaming import StreamingRunner
    runner = StreamingRunner(model, depth=2)

    def e2e_run(k, count):
        d2h = 0
        prev = None
        runner.submit(host[0][0], host[0][1], args.blocks)
        for i in range(k):
            flush.fill_(i & 0xff)
            if i + 1 < k:
                nl, nf = host[(i + 1) % args.sets]
                runner.submit(nl, nf, args.blocks)             # H2D of step i+1 overlaps the compute of step i
            t = runner.step(ones)
            if prev is not None:
                hl, hs = runner.result(prev)                   # host-side result of step i-1 is complete
                if count:
                    d2h += hl.numel() * hl.element_size() + hs.numel() * 4
            prev = t
        hl, hs = runner.result(prev)
        if count:
            d2h += hl.numel() * hl.element_size() + hs.numel() * 4
        return d2h

    e2e_run(8, False)        # warm-up: every input set twice (staging buffers, pinned result buffers and the allocator settle)
    barrier()
    e0.record()
    d2h = e2e_run(args.steps, True)
    e1.record()
    barrier()
    e2e_ms = shard.max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_value = total_vox / (e2e_ms * 1e-3)
    h2d = sum(host[i % args.sets][0].numel() * host[i % args.sets][0].element_size() + host[i % args.sets][1].numel() * 4
              for i in range(args.steps))
    my_e2e_ms = e0.elapsed_time(e1)
    pcie = {'rank': rank, 'h2d_GBps': h2d / (my_e2e_ms * 1e-3) / 1e9, 'd2h_GBps': d2h / (my_e2e_ms * 1e-3) / 1e9, 'cores': bound}
    per_rank = [pcie]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, pcie)
        per_rank = gathered

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sampler.stop_flag = True
    sampler.join(2)

    # ---- roofline of the dominant kernel (conv_gather_f32_kernel, all shapes): algorithmic bytes / measured time
    import ctypes
    _t = ctypes.c_double(0)
    lib.sgnn_debug_ffma_peak(5000, ctypes.byref(_t), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    ffma_meas = _t.value
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    per_set_bytes, per_set_flops = 0.0, 0.0
    for rec in ledger:
        r = rec['R']
        rec['bytes'] = (rec['n_in'] * rec['cin'] + rec['n_out'] * rec['cout']) * 4 + 8 * r + rec['K'] * rec['cin'] * rec['cout'] * 4
        rec['flops'] = 2.0 * r * rec['cin'] * rec['cout']
        per_set_bytes += rec['bytes']
        per_set_flops += rec['flops']
    n_conv = max(n_conv, 1)
    convs_per_step = n_conv / max(args.steps, 1)
    # DRAM bytes per launch of the same kernels from the committed ncu pass (profiles/conv_dram_traffic.json, made by
    # profiles/summarize.py dram from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`): bench.py cannot run
    # under ncu itself.  Cold cache (ncu flushes L2 per launch), conv_mode tc32.
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'conv_dram_traffic.json')))
        stamp = kernel_source_hash()
        if args.conv_mode == 'tc32' and args.blocks == 32 and tj.get('kernel_source_sha256') == stamp:
            traffic, traffic_src = tj['dram_bytes_per_launch'], 'profiles/conv_dram_traffic.json: ' + tj['source']
        else:
            traffic_src = ('not reported: profiles/conv_dram_traffic.json was captured for other kernel sources / another '
                           'workload (stamp %s, now %s)' % (str(tj.get('kernel_source_sha256'))[:12], stamp[:12]))
    except Exception:
        pass
    # ledger is of set 0; all sets are statistically alike (same generator) -> scale by launches
    alg_bytes_total = per_set_bytes * args.steps
    achieved = alg_bytes_total / (conv_ms * 1e-3) / 1e9 if conv_ms > 0 else 0.0
    roofline = {
        'kernel': ('all %d convolutions per step: conv_ur_kernel<Q> + conv_urc_kernel (tcgen05, distinct rows of a tile staged once by '
                   'TMA) for the Cout=16 layers of site sets >= 1000 rows, conv_tc32_kernel for large stride-2 layers, '
                   'conv_ro_kernel (FFMA) for the rest' if args.conv_mode == 'tc32' else
                   'sgnn_conv_forward = conv_ro_kernel<COUT,CIN,..> + conv_child_f32_kernel (all %d launches per step)')
                  % round(convs_per_step),
        'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
        'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
        # the per-layer ledger the totals come from: [op, n_in, n_out, cin, cout, K, R, bytes, flops] per launch
        'ledger_set0': [[r['op'], r['n_in'], r['n_out'], r['cin'], r['cout'], r['K'], r['R'], r['bytes'], r['flops']]
                        for r in ledger],
        'algorithmic_bytes_per_launch': per_set_bytes / max(convs_per_step, 1),
        'avg_launch_us': 1e3 * conv_ms / n_conv,
        'share_of_step': conv_ms / prof_ms if prof_ms > 0 else None,
        'profiled_pass_ms_per_step': prof_ms / max(args.steps, 1),
        'phases_us_one_pass': phase_table,
        'gflops_per_step': per_set_flops / 1e9,
        'achieved_tflops_fp32': per_set_flops * args.steps / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0,
        'fp32_ffma_peak_tflops': 148 * 128 * 2 * 1.965e9 / 1e12,
        'fp32_ffma_peak_tflops_measured': ffma_meas,
        'note': ('tc32: fp32 features on tcgen05 via an exact 3-way bf16 split; activations are L2 resident (DRAM traffic below the '
                 'algorithmic bytes); the unique-row kernels are bound by per-work-item synchronisation latency (~870 cycles per 3 '
                 'filter offsets on both the MMA-issue and the producer side, profiles/r02_ur_pipeline_analysis.txt), shared-memory '
                 'data pipe 65 %, tensor pipe 17 %' if args.conv_mode == 'tc32' else
                 'fp32 FFMA path: activations are L2 resident and the kernel is FFMA / L2-gather bound, so the HBM '
                 'fraction (SURVEY 8(d) definition) is small by construction; achieved_tflops_fp32 vs the FFMA peak '
                 'is the meaningful ceiling for this dtype'),
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, t, cores, nvox = cpu_forward_rate(args.cpu_blocks if args.cpu_blocks > 0 else 8, args.cpu_reps, args.param_seed)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': 'median of %d forwards of %d x 64^3 block(s) of the same workload (%d voxels, %.2f s each, 1 warm-up); '
                         'restated SparseConvNet-CPU path (oracle O2) under the oracle port of model.py' %
                         (args.cpu_reps, args.cpu_blocks if args.cpu_blocks > 0 else 8, nvox, t)}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[%d]: %d synthetic 64^3 TSDF blocks @5%% per GPU per step, full '
                               '3-level coarse-to-fine generator, fp32' % (args.config, args.blocks),
                   'blocks_per_gpu': args.blocks, 'input_voxels_per_step_per_gpu': vox[0],
                   'level_candidates_set0': level_rows, 'output_voxels_set0': final_rows,
                   'parallelism': 'independent blocks, rank = block mod %d, one NCCL weight broadcast' % world,
                   'l2': '256 MiB flush before every step (inside the timed bracket); %d rotating input sets'
                         % args.sets,
                   'params': 'deterministic hash fill seed %d (643735 params)' % args.param_seed,
                   'conv_mode': args.conv_mode},
        'roofline': roofline, 'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d // args.steps,
                'd2h_bytes_per_step': d2h // max(args.steps, 1), 'ms_per_step': e2e_ms / args.steps,
                'coordinates': 'int16 x 4 over PCIe both ways (StreamingRunner compact mode)', 'per_rank': per_rank},
        'gpu_launches': launches, 'clocks': sampler.summary(), 'wall_s': wall,
    }
    if args.ledger:
        with open(args.ledger, 'w') as f:
            json.dump({'ledger_set0': ledger, 'bytes': per_set_bytes, 'flops': per_set_flops,
                       'conv_times_last_profiled_step': conv_table}, f, indent=1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 from C (NCCL's version banner under
    # torchrun, for one) are sent to stderr; the line itself goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, 'w', buffering=1)
    if args.config in (2, 4):
        if int(os.environ.get('RANK', '0')) == 0:
            import bench_ops
            bench_ops.contract_line(args)
    elif args.impl == 'reference':
        if args.blocks <= 0:
            args.blocks = 32          # the CPU arm samples configs[3] with the configs[1] step (same per-voxel work)
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == '__main__':
    main()

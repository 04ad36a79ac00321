"""bench.py -- BASELINE.json metric: Mvoxels/s of the sparse U-Net forward (+ offset clustering) per tile.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2_2M] [--mode fp32|tf32]
  python bench.py --impl reference ...        (the reference's CPU path = oracle port, host cores)

A "step" = one pass of the hot path over one synthetic forest tile per GPU: point->voxel, level pyramid +
rulebooks, every sparse conv of the 7-level U-Net, voxel->point gather + heads, then the offset-shifted
clustering (DBSCAN-equivalent) and kNN assignment of the remaining tree points.
  value = level-0 active voxels of all ranks / max-over-ranks device time, inputs resident in HBM.
  e2e   = same metric through the public per-tile call `treelearn_b200.pipeline.segment_tile` with the
          batch in pinned HOST memory (H2D of coords/feats/batch ids and D2H of the labels inside the timing).
Multi-GPU: tiles shard one per GPU with no data-path collective => "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GROUPING = SimpleNamespace(tree_conf_thresh=0.5, tau_vert=0.6, tau_off=4, tau_group=0.15, tau_min=50, use_hdbscan=False)
SPATIAL_SHAPE = [1000, 1000, 1000]     # the 60 m cfg-2 tile exceeds the default [500,500,1000] in xy
MODEL_CFG = dict(channels=32, num_blocks=7, use_feats=False, use_coords=False, spatial_shape=SPATIAL_SHAPE)
METRIC = 'Mvoxels/s sparse U-Net fwd (+cluster) per tile'


def conv_traffic(workload, mode):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of all k_conv_tc launches of one forward, from the
    committed ncu capture of the same workload / mode (profiles/r01_conv_traffic.json, written by
    tools/summarise_conv_traffic.py); None when no capture matches."""
    path = os.path.join(ROOT, 'profiles', 'r01_conv_traffic.json')
    if not os.path.exists(path):
        return None
    t = json.load(open(path))
    return t.get('dram_bytes_per_step') if (t.get('workload'), t.get('mode')) == (workload, mode) else None


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


def make_tile(workload, rank):
    from treelearn_b200 import synth
    cfg = dict(synth.WORKLOADS[workload])
    cfg['seed'] = cfg['seed'] + 100 * rank
    return synth.make_batch([synth.synth_forest(**cfg)])


def plot_record(net, args, world, rank, dev, dist):
    """BASELINE.json config 4 -- whole-plot inference, STRONG scaling: one synthetic plot cut into 64 overlapping 35 m tiles
    (tools/pipeline/pipeline.py:62-94 of the reference), tiles sharded over the ranks, one all-gather of the inner rows,
    replicated overlap merge + clustering + kNN assignment, instance labels on every rank.  Times the public call
    `treelearn_b200.dist.segment_plot` with the tile batches in pinned HOST memory (H2D inside), max over ranks; then rank 0
    repeats the plot alone for the 1-GPU time the speed-up is quoted against."""
    from treelearn_b200 import dist as tdist, synth
    tiles = synth.plot_tiles(n_side=8)
    tiles = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in t.items()} for t in tiles]
    n_points = sum(int(t['coords'].shape[0]) for t in tiles)

    def once(group_world):
        marks = []
        if group_world > 1:
            coords, labels, ncl = tdist.segment_plot(net, tiles, GROUPING, marks=marks)
        else:   # single rank: every tile here, no collective
            saved = (dist.is_initialized, dist.get_world_size, dist.get_rank)
            dist.is_initialized, dist.get_world_size, dist.get_rank = (lambda: False), (lambda g=None: 1), (lambda g=None: 0)
            try:
                coords, labels, ncl = tdist.segment_plot(net, tiles, GROUPING, marks=marks)
            finally:
                dist.is_initialized, dist.get_world_size, dist.get_rank = saved
        torch.cuda.synchronize()
        ev = dict(marks)
        names = ['forward', 'allgather', 'merge', 'cluster']
        prev, out = ev['start'], {}
        for nme in names:
            out[nme] = prev.elapsed_time(ev[nme])
            prev = ev[nme]
        out['total'] = ev['start'].elapsed_time(ev['cluster'])
        return out, int(coords.shape[0]), int(ncl)

    def timed(group_world, reps):
        once(group_world)                                    # warm-up (allocator, NCCL channels)
        best = None
        for _ in range(reps):
            if world > 1 and group_world > 1:
                dist.barrier()
            t, npts, ncl = once(group_world)
            v = torch.tensor([t[k] for k in ('forward', 'allgather', 'merge', 'cluster', 'total')], device=dev)
            if world > 1 and group_world > 1:
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
            v = v.tolist()
            if best is None or v[4] < best[0][4]:
                best = (v, npts, ncl)
        return best

    (fwd, ag, mg, cl, tot), npts, ncl = timed(world, 2)
    rec = {'workload': 'cfg4_plot64: one synthetic plot (63 m edge) cut into 64 overlapping 35 m tiles (inner 8 m, stride 0.5), '
                       'tiles sharded over the ranks, all-gather of the inner rows, replicated merge + DBSCAN-equivalent '
                       'clustering + kNN assignment', 'scaling': 'strong', 'tiles': len(tiles), 'tile_points_total': n_points,
           'merged_points': npts, 'clusters': ncl, 'n_gpus': world, 's_end_to_end': round(tot / 1e3, 4),
           'forward_ms': round(fwd, 2), 'allgather_ms': round(ag, 2), 'merge_ms': round(mg, 2), 'cluster_knn_ms': round(cl, 2),
           'Mpoints_per_s': round(n_points / (tot * 1e-3) / 1e6, 2), 'api': 'treelearn_b200.dist.segment_plot(model, tiles, grouping_cfg)'}
    if world > 1:
        if rank == 0:
            (f1, a1, m1, c1, t1), _, _ = timed(1, 1)
            rec['s_end_to_end_1gpu'] = round(t1 / 1e3, 4)
            rec['speedup_vs_1gpu'] = round(t1 / tot, 3)
            rec['serial_tail_ms_1gpu'] = round(m1 + c1, 2)
        dist.barrier()
    return rec


# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from treelearn_b200 import TreeLearn, synth, sparse, pipeline, _lib
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _lib.load()
    dev = torch.device('cuda', local)

    batch = make_tile(args.workload, rank)
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()
            if k in ('coords', 'input_feats', 'batch_ids', 'batch_size')}
    resident = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    torch.manual_seed(0)
    net = synth.randomize_bn_stats(TreeLearn(mode=args.mode, **MODEL_CFG), seed=0).to(dev).eval()
    vert_dev = resident['input_feats'][:, -1].contiguous()

    def step_resident():
        with torch.no_grad():
            out = net(resident, return_loss=False)
            labels, ncl = pipeline.instances_cuda(resident['coords'], out['offset_predictions'],
                                                  out['semantic_prediction_logits'], vert_dev, GROUPING)
        return labels, ncl

    def step_e2e():
        return pipeline.segment_tile(net, host, GROUPING)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    # level-0 voxel count of this rank's tile (one untimed pass; also the first warm-up)
    with torch.no_grad():
        _, vc, _, _ = sparse.voxelize(resident['coords'], resident['input_feats'], resident['batch_ids'], 1, 0.1, False,
                                      False, 3)
    n_vox = torch.tensor([vc.shape[0]], device=dev, dtype=torch.float64)
    n_pts = int(resident['coords'].shape[0])
    if world > 1:
        dist.all_reduce(n_vox)
    n_vox_total = n_vox.item()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sparse.PROFILE = []
    lib.tl_reset_launch_count()
    ms_res = timed(step_resident, args.steps, args.warmup)
    launches = lib.tl_launch_count() // (args.steps + args.warmup)
    prof = sparse.PROFILE
    sparse.PROFILE = None
    # split the step: backbone only (fwd) vs fwd+cluster
    def fwd_only():
        with torch.no_grad():
            net(resident, return_loss=False)
    ms_fwd = timed(fwd_only, args.steps, 1)
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.summary() if rank == 0 else None

    # roofline of the dominant kernel (the segmented gather-GEMM conv), from live CUDA events
    per_step = len(prof) // (args.steps + args.warmup)
    timed_prof = prof[args.warmup * per_step:]
    conv_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in timed_prof)
    conv_bytes = sum(p[2] for p in timed_prof)
    conv_flops = sum(p[3] for p in timed_prof)
    peak, peak_src = peaks()
    achieved = conv_bytes / (conv_ms * 1e-3) / 1e9 if conv_ms > 0 else 0.0
    roofline = {'bound': 'hbm', 'kernel': 'k_conv_simt (segmented gather-GEMM sparse conv)' if args.mode == 'fp32'
                else f'k_conv_tc (tcgen05 {args.mode} gather-GEMM sparse conv)',
                'achieved': round(achieved, 1), 'peak': peak, 'unit': 'GB/s', 'frac': round(achieved / peak, 4),
                'traffic': conv_traffic(args.workload, args.mode), 'peak_source': peak_src, 'launches_per_step': per_step,
                'kernel_ms_per_step': round(conv_ms / args.steps, 3),
                'kernel_share_of_step': round(conv_ms / args.steps / ms_res, 3),
                'alg_bytes_per_step': int(conv_bytes / args.steps),
                'dense_tflops': round(conv_flops / (conv_ms * 1e-3) / 1e12, 2) if conv_ms > 0 else 0.0}

    plot = None if args.no_plot else plot_record(net, args, world, rank, dev, dist)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = n_vox_total / (ms_res * 1e-3) / 1e6
    h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))
    line = {
        'metric': METRIC, 'value': round(value, 2), 'unit': 'Mvoxels/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms_res, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': {'fp32': 'f32', 'tf32': 'tf32 (tcgen05 kind::tf32, fp32 accumulate; fp32 storage)', 'f16': 'f16 operands (tcgen05 kind::f16, fp32 accumulate, fp32 residual stream)'}[args.mode], 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: one synthetic forest tile per GPU, {int(n_vox_total / world)} active '
                               f'0.1 m voxels, default 7-level 32-channel U-Net (random init, BN eval, randomised stats) '
                               f'+ DBSCAN-equivalent clustering + kNN assignment', 'spatial_shape': SPATIAL_SHAPE,
                   'points_per_tile': n_pts, 'mode': args.mode,
                   'l2': 'per-step working set (GBs of feature maps + rulebooks) exceeds the 126 MB L2; no explicit flush'},
        'fwd_only': {'ms_per_step': round(ms_fwd, 3), 'value': round(n_vox_total / (ms_fwd * 1e-3) / 1e6, 2),
                     'unit': 'Mvoxels/s'},
        'e2e': {'value': round(n_vox_total / (ms_e2e * 1e-3) / 1e6, 2), 'unit': 'Mvoxels/s',
                'ms_per_step': round(ms_e2e, 3), 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(n_pts * 4),
                'api': 'treelearn_b200.pipeline.segment_tile(model, host_batch, grouping_cfg)'},
        'gpu_launches': int(launches * args.steps), 'roofline': roofline, 'clocks': clocks,
    }
    if plot is not None:
        line['plot'] = plot
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(budget_s=20.0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
def cpu_step(sd, batch):
    """One pass of the reference's CPU path restated by the oracle: U-Net forward + clustering + kNN."""
    from oracle import cluster_ref, model_ref
    with torch.no_grad():
        out = model_ref.forward_ref(sd, batch, spatial_shape=SPATIAL_SHAPE)
    coords = batch['coords'].numpy()
    offs = out['offset_predictions'].numpy()
    inst = cluster_ref.get_instances_ref(coords, offs, out['semantic_prediction_logits'].numpy(), 0.5, 0.6, 4, 0.15, 50,
                                         batch['input_feats'].numpy()[:, -1])
    tm = inst != 0
    if (inst[tm] == -1).any() and (inst[tm] != -1).sum() >= 5:
        inst[tm] = cluster_ref.assign_remaining_ref(coords[tm] + offs[tm], inst[tm], -1)
    return out['backbone_feats'].shape[0]


def cpu_sample():
    from oracle import model_ref
    from treelearn_b200 import synth
    tile = synth.synth_forest(edge=14.0, n_trees=12, seed=1, ground_density=1000.0)   # same generator, bounded sample
    batch = synth.make_batch([tile])
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=0)
    return sd, batch, 'synthetic forest tile edge=14 m (same generator/density as cfg2_2M), full default U-Net + clustering'


CPU_THREADS = min(os.cpu_count() or 1, 32)   # the oracle's small GEMMs slow down beyond ~32 threads (measured on the 128-core box)


def cpu_baseline(budget_s=20.0):
    torch.set_num_threads(CPU_THREADS)
    sd, batch, desc = cpu_sample()
    n_vox = len(np.unique((np.floor((batch['coords'].numpy() - batch['coords'].numpy().min(0)) / np.float32(0.1))
                           ).astype(np.int64), axis=0))
    t0 = time.time()
    reps = 0
    while reps < 1 or (time.time() - t0 < budget_s and reps < 5):
        cpu_step(sd, batch)
        reps += 1
    dt = (time.time() - t0) / reps
    return {'value': round(n_vox / dt / 1e6, 4), 'unit': 'Mvoxels/s', 'cores': CPU_THREADS, 'kind': 'port',
            'sample': f'{desc}; {n_vox} voxels, {reps} runs, {dt:.2f} s each (oracle: torch CPU fp32 + numpy/scipy)'}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation cannot be installed (spconv absent, no network), so the
    oracle port of it is timed on the host cores.  Rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    torch.set_num_threads(CPU_THREADS)
    sd, batch, desc = cpu_sample()
    n_vox = len(np.unique((np.floor((batch['coords'].numpy() - batch['coords'].numpy().min(0)) / np.float32(0.1))
                           ).astype(np.int64), axis=0))
    for _ in range(min(args.warmup, 1)):
        cpu_step(sd, batch)
    t0 = time.time()
    for _ in range(args.steps):
        cpu_step(sd, batch)
    dt = (time.time() - t0) / args.steps
    v = round(n_vox / dt / 1e6, 4)
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'Mvoxels/s',
            'n_gpus': int(os.environ.get('WORLD_SIZE', args.gpus)), 'steps': args.steps, 'warmup': min(args.warmup, 1),
            'ms_per_step': round(dt * 1e3, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'bounded sample of {args.workload}: {desc}', 'spatial_shape': SPATIAL_SHAPE},
            'cpu_baseline': {'value': v, 'unit': 'Mvoxels/s', 'cores': CPU_THREADS, 'kind': 'port',
                             'sample': f'{desc}; {n_vox} voxels per step'},
            'e2e': {'value': v, 'unit': 'Mvoxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg2_2M')
    ap.add_argument('--mode', default='f16', choices=['fp32', 'tf32', 'f16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-plot', action='store_true', help='skip the cfg-4 whole-plot (strong scaling) record')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)

"""bench.py -- BASELINE.json metric: Mvoxels/s of the sparse U-Net forward (+ offset clustering) per tile.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2_2M] [--mode f16x2|f16|tf32|fp32]
  python bench.py --impl reference ...        (the reference's CPU path = oracle port, host cores)

A "step" = one pass of the hot path over one synthetic forest tile per GPU: point->voxel, level pyramid +
rulebooks, every sparse conv of the 7-level U-Net, voxel->point gather + heads, then the offset-shifted
clustering (DBSCAN-equivalent) and kNN assignment of the remaining tree points.
  value = level-0 active voxels of all ranks / max-over-ranks device time, inputs resident in HBM.
  e2e   = same metric through the public per-tile call `treelearn_b200.pipeline.segment_tile` with the
          batch in pinned HOST memory (H2D of coords/feats/batch ids and D2H of the labels inside the timing).
Multi-GPU: tiles shard one per GPU with no data-path collective => "scaling": "weak".

The headline mode is f16x2 (two fp16 terms per operand, three tcgen05 MMAs per K step): the one tensor-core mode whose
per-point offsets stay within BASELINE's 1e-3 of the fp32 reference at metre-scale outputs (`parity` record, measured on
BASELINE config 1 against the oracle inside the cpu_baseline leg).  The single-term mode f16 is reported beside it as
`fast_mode` with its own error.  The heads' last Linear layers are probe-fitted and the outputs moved towards the labels
(synth.fit_probe_heads, synth.TrainedLikeOutputs) so that the clustering / kNN stages see what they see behind a trained
network; `cluster_trained_like` times those stages alone, DBSCAN-equivalent and HDBSCAN.
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
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of all sparse-conv launches of one forward, from the
    committed ncu capture of the same workload / mode (profiles/r02_conv_traffic.json, written by
    tools/summarise_conv_traffic.py).  Returns (bytes or None, note): None when no capture matches the workload, the
    mode and the current conv kernel sources (a stale capture is not reported as this kernel's traffic)."""
    from treelearn_b200._lib import conv_source_hash
    path = os.path.join(ROOT, 'profiles', 'r02_conv_traffic.json')
    if not os.path.exists(path):
        return None, 'no capture committed'
    caps = json.load(open(path))
    for c in (caps if isinstance(caps, list) else [caps]):
        if (c.get('workload'), c.get('mode')) == (workload, mode):
            if c.get('conv_source_sha1') == conv_source_hash():
                return c.get('dram_bytes_per_step'), 'profiles/r02_conv_traffic.json (same kernel sources)'
            return None, f"stale: capture of sources {str(c.get('conv_source_sha1'))[:10]} measured {c.get('dram_bytes_per_step')} B/step"
    return None, 'no capture for this workload/mode'


def secondary_kernels(peak):
    """Achieved DRAM GB/s of the non-conv kernels of the step (rulebook, scatter, clustering ...: the `north_star` evidence
    list) from the committed ncu launch list profiles/r02_ncu_launches_summary_step_f16x2.txt (tools/summarise_launches.py
    over `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` of tools/profile_step.py, 3 steps).
    ncu replays every launch cold-cache and serialised, so these are per-kernel DRAM rates, not shares of the live step."""
    path = os.path.join(ROOT, 'profiles', 'r02_ncu_launches_summary_step_f16x2.txt')
    if not os.path.exists(path):
        return None
    want = ('k_subm_probe', 'k_hash_build', 'k_halo_build', 'k_level_emit', 'k_emit_voxels', 'k_heads', 'k_cc_link',
            'k_knn_vote', 'k_conv_in4', 'DeviceRadixSortOnesweep')
    out = []
    for line in open(path):
        f = line.split()
        if line.startswith('#') or len(f) < 8:
            continue
        name = ' '.join(f[7:])
        hit = [w for w in want if w in name]
        if not hit or any(o['kernel'] == hit[0] for o in out):
            continue
        us, n, rd, wr, gbs = float(f[1]), int(f[2]), float(f[3]), float(f[4]), float(f[5])
        out.append({'kernel': hit[0], 'launches_per_step': round(n / 3, 1), 'us_per_step': round(us / 3, 1),
                    'dram_MB_per_step': round((rd + wr) / 3, 1), 'achieved_GBs': gbs, 'frac': round(gbs / peak, 4)})
    return {'source': 'profiles/r02_ncu_launches_summary_step_f16x2.txt (ncu, cold-cache, serialised; 3 steps)', 'bound': 'hbm',
            'kernels': out}


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
    net = synth.TrainedLikeOutputs(net).eval()          # trained-like outputs, so that merge + clustering see real trees
    for t in tiles:
        net.prepare(t)

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
                       'clustering + kNN assignment; a rank collates up to 6 M points (8 tiles) into one network forward',
           'scaling': 'strong', 'tiles': len(tiles), 'tile_points_total': n_points,
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


def train_record(args, world, rank, dev, dist):
    """BASELINE.json configs 3 and 5 -- the training step of tools/training/train.py:32-44 (forward + loss under fp16
    autocast, GradScaler, backward, gradient clip, AdamW) through `treelearn_b200.dist.train_step`, batches in pinned HOST
    memory.  cfg3 (1 GPU only): one 4-tile batch.  cfg5: data parallel, 2 tiles per GPU, level-bucketed NCCL all-reduce
    launched from backward hooks (weak scaling: voxels/s summed over the ranks, time = max over ranks)."""
    from treelearn_b200 import TreeLearn, synth, sparse
    from treelearn_b200 import dist as tdist

    def run(n_tiles, seed0, steps=8, warmup=6):   # GradScaler skips its first ~4 steps (scale 65536 -> 4096); AdamW state is created by the first real one
        tiles = [synth.synth_forest(edge=20.0, n_trees=20, seed=seed0 + s) for s in range(n_tiles)]
        batch = synth.make_batch(tiles)
        batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
        with torch.no_grad():
            _, vc, _, _ = sparse.voxelize(batch['coords'].to(dev), batch['input_feats'].to(dev), batch['batch_ids'].to(dev),
                                          n_tiles, 0.1, False, False, 3)
        n_vox = torch.tensor([float(vc.shape[0])], device=dev, dtype=torch.float64)
        torch.manual_seed(0)
        net = TreeLearn(mode='tf32', **MODEL_CFG).to(dev).train()
        opt = torch.optim.AdamW(net.parameters(), lr=2e-3, weight_decay=1e-3)
        scaler = torch.amp.GradScaler('cuda')
        red = tdist.OverlappedGradReducer(net)
        exposed = []
        finish = red.finish

        def timed_finish():            # device time between the end of backward and the end of the last all-reduce
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            finish()
            b.record()
            exposed.append((a, b))
        red.finish = timed_finish
        for _ in range(warmup):
            tdist.train_step(net, opt, batch, reducer=red, scaler=scaler, autocast=True, grad_clip=10.0)
        exposed.clear()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [e0]
        e0.record()
        for _ in range(steps):
            loss, _ = tdist.train_step(net, opt, batch, reducer=red, scaler=scaler, autocast=True, grad_clip=10.0)
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        e1.record()
        torch.cuda.synchronize()
        per_step = [round(a.elapsed_time(b), 1) for a, b in zip(marks[:-1], marks[1:])]
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps, sum(a.elapsed_time(b) for a, b in exposed) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(n_vox)
        red.remove()
        ms_step, ms_exposed = ms.tolist()
        out = {'tiles_per_gpu': n_tiles, 'voxels_total': int(n_vox.item()), 'ms_per_step': round(ms_step, 2),
               'Mvoxels_per_s': round(n_vox.item() / (ms_step * 1e-3) / 1e6, 2), 'loss': round(float(loss.item()), 4),
               'steps': steps, 'warmup': warmup, 'ms_each_step': per_step, 'grad_scale': float(scaler.get_scale()),
               'peak_mem_GiB': round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
        if world > 1:
            out['allreduce_exposed_ms'] = round(ms_exposed, 2)
        del net, opt
        torch.cuda.empty_cache()
        return out

    rec = {'step': 'treelearn_b200.dist.train_step: fp16 autocast + GradScaler (tools/training/train.py:32-44), TF32 tcgen05 '
                   'conv forward / data-gradient kernels, TF32 weight gradients, batch-statistics BatchNorm, clip 10, AdamW; '
                   'random-init default U-Net, 20 m synthetic tiles, host batches',
           'n_gpus': world}
    if world == 1:
        rec['cfg3_4tile_batch'] = run(4, 0)
    rec['cfg5_dp_2tiles_per_gpu'] = dict(run(2, 10 * rank), scaling='weak',
                                         allreduce='one NCCL all-reduce per U-Net level bucket, launched from backward hooks')
    return rec


# ----------------------------------------------------------------------------------------------------
DTYPE_TEXT = {
    'fp32': 'f32',
    'tf32': 'tf32 (tcgen05 kind::tf32, fp32 accumulate; fp32 storage)',
    'f16': 'f16 operands (tcgen05 kind::f16, fp32 accumulate, fp32 residual stream)',
    'f16x2': 'f16x2: two fp16 terms per operand (hi + lo), 3 tcgen05 kind::f16 MMAs per K step, fp32 accumulate, fp32 '
             'residual stream (~2^-21 relative operand error)',
    'mixed': 'mixed: f16x2 (two fp16 terms per operand) on U-Net levels 0-1, f16 (one term) on levels 2-6; fp32 accumulate, fp32 '
             'residual stream',
}
CONV_KERNEL = {
    'fp32': 'k_conv_simt (segmented gather-GEMM sparse conv, fp32 FMA)',
    'tf32': 'k_conv_tc (tcgen05 tf32 gather-GEMM sparse conv)',
    'f16': 'k_conv_halo<1> + k_conv_grp<1> (tcgen05 f16 sparse conv: halo-cached TS-form kernel for the 3^3 submanifold layers, '
           'gather kernel for the strided / inverse layers); all conv launches of the step',
    'mixed': 'k_conv_halo<2|1> + k_conv_grp<2|1> (tcgen05 f16 sparse conv, two-term operands on levels 0-1); all conv launches of the step',
    'f16x2': 'k_conv_halo<2> + k_conv_grp<2> (tcgen05 f16 hi/lo sparse conv: halo-cached TS-form kernel for the 3^3 submanifold '
             'layers, gather kernel for the strided / inverse layers); all conv launches of the step',
}


def build_net(mode, dev, world, dist):
    """Default U-Net, random init, BN eval with randomised statistics, probe-fitted last Linear of both heads (fitted on
    the cfg-1 tile in fp32 on rank 0 and broadcast, so every rank holds the same weights)."""
    from treelearn_b200 import TreeLearn, synth
    torch.manual_seed(0)
    net = synth.randomize_bn_stats(TreeLearn(mode='fp32', **MODEL_CFG), seed=0).to(dev).eval()
    fit_tile = synth.make_batch([synth.workload('cfg1_200k')])
    synth.fit_probe_heads(net, fit_tile)
    if world > 1:
        for p in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(p.data, 0)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}

    def make(m):
        n = TreeLearn(mode=m, **MODEL_CFG)
        n.load_state_dict(sd)
        return n.to(dev).eval()
    return make, {k: v.cpu() for k, v in sd.items()}, fit_tile


def cluster_record(args, resident, dev):
    """Offset-shifted clustering + kNN assignment alone, on trained-like predictions of the bench tile (labels + 5 cm
    Gaussian noise on tree points): the DBSCAN-equivalent branch on the whole tile, and the HDBSCAN branch on the points of
    the tile's central 20 m x 20 m (the GPU Prim MST is O(n^2))."""
    from treelearn_b200 import pipeline
    g = torch.Generator(device='cpu').manual_seed(5)
    labels = resident['offset_labels']
    tree = resident['semantic_labels'] == 0
    noise = (0.05 * torch.randn(labels.shape, generator=g)).to(dev)
    offs = torch.where(tree[:, None], labels + noise, torch.zeros(1, device=dev)).contiguous()
    logits = (torch.where(tree, 4.0, -4.0)[:, None] * torch.tensor([1.0, -1.0], device=dev)).contiguous()
    vert = resident['input_feats'][:, -1].contiguous()
    coords = resident['coords']

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    ms_d, (_, ncl_d) = timed(lambda: pipeline.instances_cuda(coords, offs, logits, vert, GROUPING), 3)
    rec = {'predictions': 'offset labels + N(0, 5 cm) on tree points, logits +-4 by label (trained-like)',
           'dbscan': {'points': int(coords.shape[0]), 'clusters': int(ncl_d), 'ms': round(ms_d, 3),
                      'Mpoints_per_s': round(coords.shape[0] / (ms_d * 1e-3) / 1e6, 2)}}
    sel = (coords[:, 0].abs() <= 10) & (coords[:, 1].abs() <= 10)
    hcfg = SimpleNamespace(**{**vars(GROUPING), 'use_hdbscan': True})
    c2, o2, l2, v2 = coords[sel].contiguous(), offs[sel].contiguous(), logits[sel].contiguous(), vert[sel].contiguous()
    n_filtered = int(((l2[:, 0] > 0) & (v2 > GROUPING.tau_vert) & (o2[:, 2].abs() < GROUPING.tau_off)).sum())
    ms_h, (_, ncl_h) = timed(lambda: pipeline.instances_cuda(c2, o2, l2, v2, hcfg), 1)
    rec['hdbscan'] = {'points': int(c2.shape[0]), 'clustered_points': n_filtered, 'clusters': int(ncl_h), 'ms': round(ms_h, 3),
                      'Mpoints_per_s': round(c2.shape[0] / (ms_h * 1e-3) / 1e6, 3),
                      'sample': 'central 20 m x 20 m of the tile (GPU Prim MST is O(n^2) in the clustered points)'}
    return rec


def run_b200(args):
    import torch.distributed as dist
    from treelearn_b200 import synth, sparse, pipeline, _lib
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _lib.load()
    dev = torch.device('cuda', local)

    batch = make_tile(args.workload, rank)
    in_keys = ('coords', 'input_feats', 'batch_ids', 'batch_size')
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items() if k in in_keys}
    resident = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()
                if k in in_keys + ('offset_labels', 'semantic_labels')}
    make_net, sd_cpu, fit_tile = build_net(args.mode, dev, world, dist)
    raw_net = make_net(args.mode)
    net = synth.TrainedLikeOutputs(raw_net).eval()
    net.prepare(resident)
    net.share(host, resident)
    vert_dev = resident['input_feats'][:, -1].contiguous()

    def step_resident(model=net):
        with torch.no_grad():
            out = model(resident, return_loss=False)
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
    _, n_clusters = step_resident()

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
            raw_net(resident, return_loss=False)
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
    traffic, traffic_note = conv_traffic(args.workload, args.mode)
    roofline = {'bound': 'hbm', 'kernel': CONV_KERNEL[args.mode],
                'achieved': round(achieved, 1), 'peak': peak, 'unit': 'GB/s', 'frac': round(achieved / peak, 4),
                'traffic': traffic, 'traffic_note': traffic_note, 'peak_source': peak_src, 'launches_per_step': per_step,
                'kernel_ms_per_step': round(conv_ms / args.steps, 3),
                'kernel_share_of_step': round(conv_ms / args.steps / ms_res, 3),
                'alg_bytes_per_step': int(conv_bytes / args.steps),
                'dense_tflops': round(conv_flops / (conv_ms * 1e-3) / 1e12, 2) if conv_ms > 0 else 0.0}
    sec = secondary_kernels(peak)
    if sec is not None:
        roofline['secondary'] = sec

    # the single-term mode beside the headline (same weights, same tile, same trained-like correction)
    fast = mixed = None
    if args.mode == 'f16x2':
        def conv_stack(prof):    # roofline of the conv launches of the timed steps of a secondary mode, like `roofline`
            per = len(prof) // (args.steps + args.warmup)
            tp = prof[args.warmup * per:]
            ms = sum(a.elapsed_time(b) for a, b, _, _, _ in tp)
            by = sum(p[2] for p in tp)
            return {'kernel_ms_per_step': round(ms / args.steps, 3), 'alg_bytes_per_step': int(by / args.steps),
                    'achieved': round(by / (ms * 1e-3) / 1e9, 1), 'unit': 'GB/s', 'frac': round(by / (ms * 1e-3) / 1e9 / peak, 4)}

        fnet = synth.TrainedLikeOutputs(make_net('f16')).eval()
        fnet._corr = net._corr
        sparse.PROFILE = []
        ms_fast = timed(lambda: step_resident(fnet), args.steps, args.warmup)
        fast_conv, sparse.PROFILE = conv_stack(sparse.PROFILE), None
        fast = {'mode': 'f16', 'dtype': DTYPE_TEXT['f16'], 'ms_per_step': round(ms_fast, 3),
                'value': round(n_vox_total / (ms_fast * 1e-3) / 1e6, 2), 'unit': 'Mvoxels/s',
                'conv_stack': fast_conv,
                'note': 'outside the 1e-3 offset tolerance at metre-scale outputs: see parity.fast_mode'}
        del fnet
        # two fp16 terms on U-Net levels 0-1 only (94 % of the voxels), one term on levels 2-6: inside the tolerance with a
        # ~5x margin (parity.mixed_mode), reported beside the all-levels f16x2 headline
        mnet = synth.TrainedLikeOutputs(make_net('mixed')).eval()
        mnet._corr = net._corr
        sparse.PROFILE = []
        ms_mixed = timed(lambda: step_resident(mnet), args.steps, args.warmup)
        mixed_conv, sparse.PROFILE = conv_stack(sparse.PROFILE), None
        mixed = {'mode': 'mixed', 'dtype': DTYPE_TEXT['mixed'], 'ms_per_step': round(ms_mixed, 3),
                 'value': round(n_vox_total / (ms_mixed * 1e-3) / 1e6, 2), 'unit': 'Mvoxels/s',
                 'conv_stack': mixed_conv, 'note': 'see parity.mixed_mode for its error against the fp32 oracle'}
        del mnet

    cluster = cluster_record(args, resident, dev) if (rank == 0 and not args.no_cluster) else None
    plot = None if args.no_plot else plot_record(raw_net, args, world, rank, dev, dist)
    train = None if args.no_train else train_record(args, world, rank, dev, dist)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = n_vox_total / (ms_res * 1e-3) / 1e6
    # segment_tile copies coords and input_feats once; batch_ids of a one-tile batch are filled on the device (model.py)
    h2d = sum(host[k].numel() * host[k].element_size() for k in ('coords', 'input_feats'))
    line = {
        'metric': METRIC, 'value': round(value, 2), 'unit': 'Mvoxels/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms_res, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': DTYPE_TEXT[args.mode], 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: one synthetic forest tile per GPU, {int(n_vox_total / world)} active '
                               f'0.1 m voxels, default 7-level 32-channel U-Net (random init, BN eval, randomised stats, '
                               f'probe-fitted heads, trained-like outputs) + DBSCAN-equivalent clustering + kNN assignment',
                   'spatial_shape': SPATIAL_SHAPE, 'points_per_tile': n_pts, 'mode': args.mode,
                   'clusters_per_tile': int(n_clusters),
                   'l2': 'per-step working set (GBs of feature maps + rulebooks) exceeds the 126 MB L2; no explicit flush'},
        'fwd_only': {'ms_per_step': round(ms_fwd, 3), 'value': round(n_vox_total / (ms_fwd * 1e-3) / 1e6, 2),
                     'unit': 'Mvoxels/s'},
        'e2e': {'value': round(n_vox_total / (ms_e2e * 1e-3) / 1e6, 2), 'unit': 'Mvoxels/s',
                'ms_per_step': round(ms_e2e, 3), 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(n_pts * 4),
                'api': 'treelearn_b200.pipeline.segment_tile(model, host_batch, grouping_cfg)'},
        'gpu_launches': int(launches * args.steps), 'roofline': roofline, 'clocks': clocks,
    }
    if mixed is not None:
        line['mixed_mode'] = mixed
    if fast is not None:
        line['fast_mode'] = fast
    if cluster is not None:
        line['cluster_trained_like'] = cluster
    if plot is not None:
        line['plot'] = plot
    if train is not None:
        line['train'] = train
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'], line['parity'] = cpu_baseline_and_parity(make_net, sd_cpu, fit_tile, args.mode, budget_s=20.0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
def cpu_forward(sd, batch, shape):
    from oracle import model_ref
    with torch.no_grad():
        return model_ref.forward_ref(sd, batch, spatial_shape=shape)


def cpu_cluster(batch, offs, logits):
    """The reference's CPU clustering + kNN assignment restated by the oracle (tools/pipeline/pipeline.py:89-94)."""
    from oracle import cluster_ref
    coords = batch['coords'].numpy()
    inst = cluster_ref.get_instances_ref(coords, offs, logits, 0.5, 0.6, 4, 0.15, 50, batch['input_feats'].numpy()[:, -1])
    tm = inst != 0
    if (inst[tm] == -1).any() and (inst[tm] != -1).sum() >= 5:
        inst[tm] = cluster_ref.assign_remaining_ref(coords[tm] + offs[tm], inst[tm], -1)
    return inst


KEEP = 0.9


def trained_like(batch, ref, out):
    """The correction of synth.TrainedLikeOutputs computed from the ORACLE's outputs and applied to `out` (numpy)."""
    tree = batch['semantic_labels'] == 0
    c_off = torch.where(tree[:, None], batch['offset_labels'] - KEEP * ref['offset_predictions'], torch.zeros(1))
    c_sem = torch.where(tree, 4.0, -4.0)[:, None] * torch.tensor([1.0, -1.0]) - KEEP * ref['semantic_prediction_logits']
    return (out['offset_predictions'].float().cpu() + c_off).numpy(), (out['semantic_prediction_logits'].float().cpu() + c_sem).numpy()


def cpu_step(sd, batch, shape=SPATIAL_SHAPE):
    """One pass of the reference's CPU path restated by the oracle: U-Net forward + clustering + kNN on trained-like
    outputs (same correction as the GPU arm's)."""
    ref = cpu_forward(sd, batch, shape)
    offs, logits = trained_like(batch, ref, ref)
    return ref, cpu_cluster(batch, offs, logits)


CPU_THREADS = min(os.cpu_count() or 1, 32)   # the oracle's small GEMMs slow down beyond ~32 threads (measured on the 128-core box)
CPU_SAMPLE = 'cfg1_200k: BASELINE config 1, one 20 m synthetic forest tile, full default U-Net + clustering + kNN'
CPU_SHAPE = [500, 500, 1000]


def cpu_sample():
    from oracle import model_ref
    from treelearn_b200 import synth
    batch = synth.make_batch([synth.workload('cfg1_200k')])
    sd = model_ref.make_state_dict(channels=32, num_blocks=7, seed=0)
    return sd, batch


def count_voxels(batch):
    c = batch['coords'].numpy()
    return len(np.unique(np.floor((c - c.min(0)) / np.float32(0.1)).astype(np.int64), axis=0))


def cpu_baseline_and_parity(make_net, sd, batch, mode, budget_s=20.0):
    """cpu_baseline: the oracle (port of the reference's CPU path) timed on BASELINE config 1 on the host cores.
    parity: the oracle outputs of that same run are the checker for the GPU path on the same tile and weights: per-point
    offsets / logits (tolerance 1e-3, BASELINE north_star) and the instances clustered from either side."""
    from oracle import post_ref
    torch.set_num_threads(CPU_THREADS)
    n_vox = count_voxels(batch)
    t0 = time.time()
    reps = 0
    while reps < 1 or (time.time() - t0 < budget_s and reps < 5):
        ref, inst_ref = cpu_step(sd, batch, CPU_SHAPE)
        reps += 1
    dt = (time.time() - t0) / reps
    base = {'value': round(n_vox / dt / 1e6, 4), 'unit': 'Mvoxels/s', 'cores': CPU_THREADS, 'kind': 'port',
            'sample': f'{CPU_SAMPLE}; {n_vox} voxels, {reps} runs, {dt:.2f} s each (oracle: torch CPU fp32 + numpy/scipy)'}

    def check(m):
        from treelearn_b200 import TreeLearn
        net = TreeLearn(mode=m, **{**MODEL_CFG, 'spatial_shape': CPU_SHAPE})
        net.load_state_dict(sd)
        net = net.cuda().eval()
        with torch.no_grad():
            out = {k: v.float().cpu() for k, v in net(batch, return_loss=False).items()}
        eo = (out['offset_predictions'] - ref['offset_predictions']).abs().max().item()
        el = (out['semantic_prediction_logits'] - ref['semantic_prediction_logits']).abs().max().item()
        offs, logits = trained_like(batch, ref, out)
        inst = cpu_cluster(batch, offs, logits)
        tree = (inst_ref > 0) | (inst > 0)
        mg, mp, iou, _, _ = post_ref.get_detections_ref(inst_ref[tree], inst[tree], 0.5, 0)
        return {'mode': m, 'offset_max_abs_err': float(f'{eo:.3e}'), 'logit_max_abs_err': float(f'{el:.3e}'),
                'offsets_within_tolerance': bool(eo < 1e-3),
                'instances_oracle': int(inst_ref.max()), 'instances_cuda': int(inst.max()), 'matched': int(len(mg)),
                'min_matched_iou': round(float(iou[mp, mg].min()), 6) if len(mg) else None,
                'point_labels_differing': int((inst != inst_ref).sum())}

    par = {'workload': CPU_SAMPLE, 'tolerance': 1e-3,
           'offset_abs_max_m': round(float(ref['offset_predictions'].abs().max()), 2),
           'checker': 'oracle/model_ref.forward_ref + oracle/cluster_ref (fp32 CPU restatement of the reference)',
           'instances': 'clustered by the oracle from either side\'s outputs after the same trained-like correction',
           'headline': check(mode)}
    if mode == 'f16x2':
        par['mixed_mode'] = check('mixed')
        par['fast_mode'] = check('f16')
    return base, par


def run_reference(args):
    """Reference arm: the reference's own CPU implementation cannot be installed (spconv absent, no network), so the
    oracle port of it is timed on the host cores, on BASELINE config 1 (its CPU-runnable case).  Rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    from treelearn_b200 import synth
    torch.set_num_threads(CPU_THREADS)
    sd, batch = cpu_sample()
    # same weights as the GPU arm would use need a GPU for the probe fit; the CPU arm's cost does not depend on them
    n_vox = count_voxels(batch)
    for _ in range(min(args.warmup, 1)):
        cpu_step(sd, batch, CPU_SHAPE)
    t0 = time.time()
    for _ in range(args.steps):
        cpu_step(sd, batch, CPU_SHAPE)
    dt = (time.time() - t0) / args.steps
    v = round(n_vox / dt / 1e6, 4)
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'Mvoxels/s',
            'n_gpus': int(os.environ.get('WORLD_SIZE', args.gpus)), 'steps': args.steps, 'warmup': min(args.warmup, 1),
            'ms_per_step': round(dt * 1e3, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'bounded sample of {args.workload}: {CPU_SAMPLE}', 'spatial_shape': CPU_SHAPE},
            'cpu_baseline': {'value': v, 'unit': 'Mvoxels/s', 'cores': CPU_THREADS, 'kind': 'port',
                             'sample': f'{CPU_SAMPLE}; {n_vox} voxels per step'},
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
    ap.add_argument('--mode', default='f16x2', choices=['fp32', 'tf32', 'f16', 'f16x2', 'mixed'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cluster', action='store_true', help='skip the trained-like clustering record')
    ap.add_argument('--no-train', action='store_true', help='skip the cfg-3 / cfg-5 training-step record')
    ap.add_argument('--no-plot', action='store_true', help='skip the cfg-4 whole-plot (strong scaling) record')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)

"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, the model keeps the
reference's state_dict layout, the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from treelearn_b200 import TreeLearn, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'treelearn_b200.h')).read()
    declared = set(re.findall(r'\b(tl_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    assert os.path.exists(_lib.LIB_PATH), 'build the extension first (python __graft_entry__.py)'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/treelearn_b200.h but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.tl_version() >= 1


def test_ctypes_signatures_have_the_declared_argument_counts():
    """Guards the hand-written ctypes table against drifting from the header (an extra / missing argument is silent in C)."""
    header = open(os.path.join(ROOT, 'include', 'treelearn_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    decls = dict(re.findall(r'\b(tl_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', header))
    assert set(decls) == set(_lib.SIGNATURES)
    for name, params in decls.items():
        params = params.strip()
        count = 0 if params in ('', 'void') else params.count(',') + 1
        assert count == len(_lib.SIGNATURES[name][1]), (name, count, len(_lib.SIGNATURES[name][1]))


def test_state_dict_layout_matches_reference():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'model_small.npz'))
    keys = [k[3:] for k in g.files if k.startswith('sd:')]
    net = TreeLearn(channels=8, num_blocks=3)
    sd = net.state_dict()
    assert list(sd.keys()) == keys
    for k in keys:
        assert tuple(sd[k].shape) == g['sd:' + k].shape, k
    net.load_state_dict({k: torch.from_numpy(g['sd:' + k]) for k in keys}, strict=True)
    full = TreeLearn()
    n_backbone = sum(p.numel() for n, p in full.named_parameters() if not n.startswith(('semantic_', 'offset_')))
    assert n_backbone == 30104576
    assert full.state_dict()['input_conv.0.weight'].shape == (32, 3, 3, 3, 4)


def test_fixed_modules_keep_bn_in_eval():
    net = TreeLearn(channels=8, num_blocks=2, fixed_modules=['unet'])
    net.train()
    assert not net.unet.blocks.block0.conv_branch._modules['0'].training
    assert net.semantic_linear[1].training
    assert all(not p.requires_grad for p in net.unet.parameters())


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    net = TreeLearn(channels=8, num_blocks=2).eval()
    batch = {'coords': torch.rand(10, 3), 'input_feats': torch.rand(10, 1), 'batch_ids': torch.zeros(10, dtype=torch.long),
             'batch_size': 1}
    with pytest.raises(Exception):
        with torch.no_grad():
            net(batch, return_loss=False)


def test_operand_format_conversions_round_trip_on_cpu():
    """The two elementwise passes at the level boundary of mode 'mixed' (sparse.split_to_half / half_to_split) against the
    reference conversions to_split / to_p / from_split."""
    import torch
    from treelearn_b200 import sparse
    x = torch.randn(7, 96, generator=torch.Generator().manual_seed(0))
    h = sparse.split_to_half(sparse.to_split(x))                    # f16x2 operand -> one fp16 term, P-layout kept
    assert h.dtype == torch.float16 and h.shape == (7, 96)
    assert torch.equal(h, sparse.to_p(x).half())                    # hi + lo re-rounds to fp16(x) exactly
    s = sparse.half_to_split(h)
    assert s.shape == (7, 192)
    assert torch.equal(sparse.from_split(s), sparse.from_p(h.float()))   # lo terms are zero


def test_bench_secondary_kernel_table_parses_the_committed_launch_list():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(root, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sec = mod.secondary_kernels(6455.6)
    names = [k['kernel'] for k in sec['kernels']]
    for want in ('k_subm_probe', 'k_halo_build', 'k_heads', 'k_cc_link', 'k_knn_vote', 'k_emit_voxels'):
        assert want in names
    assert all(0 < k['frac'] < 1 and k['us_per_step'] > 0 for k in sec['kernels'])

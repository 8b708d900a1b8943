"""CPU: the N3 kernels' REAL source (synchformer_b200/csrc/train.cu, attention_train.cu: device code AND the C-ABI launch code),
executed on the CPU SIMT emulator of tests/emu/, driven through the REAL wrappers of synchformer_b200/ops.py, checked by the very
same test functions that run on the B200 (tests/test_train_gpu.py).  What this covers before the first hardware run: indexing, barrier
placement, shared-memory layout, launch configuration, ctypes marshalling, and the numerics of the algorithms.  What it cannot cover:
anything that depends on the hardware itself (occupancy, the real shuffle / barrier primitives, memory-model races between warps that a
coroutine schedule does not interleave, nvcc code generation)."""
import os

import pytest
import torch

import test_train_gpu as G
from emu import binding


@pytest.fixture
def cuda_device(monkeypatch):
    binding.install(monkeypatch)
    with torch.enable_grad():
        yield torch.device('cpu')


test_dropout_mask_is_the_philox_oracle_bit_for_bit = G.test_dropout_mask_is_the_philox_oracle_bit_for_bit
test_transpose_bf16_pads_with_zeros = G.test_transpose_bf16_pads_with_zeros
test_colsum = G.test_colsum
test_gelu_forward_and_backward = G.test_gelu_forward_and_backward
test_layernorm_backward = G.test_layernorm_backward
test_attention_train_forward_and_backward = G.test_attention_train_forward_and_backward
test_attention_train_rejects_unsupported_shapes = G.test_attention_train_rejects_unsupported_shapes
test_sync_head_backward = G.test_sync_head_backward
test_linear_function_gradients = G.test_linear_function_gradients
test_training_step_gradients_match_oracle_on_golden_features = G.test_training_step_gradients_match_oracle_on_golden_features


def test_training_step_at_five_second_clip_shape(cuda_device, monkeypatch):
    """one step at T = 198 (the repeat-for-determinism part runs on the GPU, and here with SFB_EMU_FULL=1: 3 x 30 s of emulation)"""
    G.test_training_step_at_five_second_clip_shape(cuda_device, monkeypatch, check_determinism=os.environ.get('SFB_EMU_FULL') == '1')


@pytest.mark.skipif(os.environ.get('SFB_EMU_FULL') != '1', reason='same kernels as the tests above, only another head; SFB_EMU_FULL=1 runs it')
def test_syncability_head_trains(cuda_device, monkeypatch):
    G.test_syncability_head_trains(cuda_device, monkeypatch)


def test_emulator_detects_a_missing_barrier_partner():
    """the scheduler aborts on a barrier that not every live thread reaches; here: every kernel of the suite above ran to completion, and the
    launch counter moved, i.e. the kernels really executed in the emulator (not in a torch stand-in)"""
    lib = binding.load()
    before = lib.emu_launch_count()
    mp = pytest.MonkeyPatch()
    try:
        binding.install(mp)
        from synchformer_b200 import ops
        ops.colsum(torch.ones((130, 64)))
    finally:
        mp.undo()
    assert lib.emu_launch_count() == before + 2

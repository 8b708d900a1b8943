"""Emulator check of the opt-in fused CLS query of the time-attention kernel (SFB_TIME_CLS_FUSED=1; csrc/attention.cu,
attn_time_mma_kernel<12, true> + sfb_attention_merge_partials).  Run as a script: the switch is read once per process.

    SFB_TIME_CLS_FUSED=1 python tests/emu/time_cls_fused_check.py   -> prints OK

Compares the attention output of MotionFormer._divided_attention(mode='time') - every token row, and the CLS row that now comes from the
merged per-location softmax states - with the dense definition of DividedAttention.forward (vit_helper.py:100-158).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import pytest  # noqa: E402
import torch  # noqa: E402

from emu import binding  # noqa: E402


def main():
    assert os.environ.get('SFB_TIME_CLS_FUSED') == '1'
    mp = pytest.MonkeyPatch()
    lib = binding.install(mp)
    from synchformer_b200 import model as M
    torch.manual_seed(0)
    n, D, TOK, h, d = 2, 768, 1569, 12, 64
    qkv = (torch.randn(n * TOK, 3 * D) * 0.7).to(torch.bfloat16)
    att = torch.zeros(n * TOK, D, dtype=torch.bfloat16)
    m = M.MotionFormer.__new__(M.MotionFormer)                      # only the method is needed
    before = lib.emu_launch_count()
    M.MotionFormer._divided_attention(m, qkv, att, n, 'time')
    launches = lib.emu_launch_count() - before
    assert launches == 2, f'expected the fused kernel + the merge (2 launches), saw {launches}'
    x = qkv.float()
    q, k, v = [t.reshape(n, TOK, h, d).permute(0, 2, 1, 3) for t in x.chunk(3, -1)]
    cls = torch.softmax(q[:, :, 0:1] @ k.transpose(-1, -2) * 0.125, -1) @ v
    re = lambda t: t.reshape(n, h, 8, 196, d).permute(0, 1, 3, 2, 4)
    q_, k_, v_ = re(q[:, :, 1:]), re(k[:, :, 1:]), re(v[:, :, 1:])
    ck, cv = k[:, :, 0:1].unsqueeze(2).expand(n, h, 196, 1, d), v[:, :, 0:1].unsqueeze(2).expand(n, h, 196, 1, d)
    out = torch.softmax(q_ @ torch.cat([ck, k_], 3).transpose(-1, -2) * 0.125, -1) @ torch.cat([cv, v_], 3)
    out = out.permute(0, 1, 3, 2, 4).reshape(n, h, 1568, d)
    ref = torch.cat([cls, out], 2).permute(0, 2, 1, 3).reshape(n * TOK, D)
    err = float((att.float() - ref).abs().max())
    rows0 = torch.arange(n) * TOK
    err_cls = float((att.float()[rows0] - ref[rows0]).abs().max())
    assert err < 1.6e-2 and err_cls < 8e-3, (err, err_cls)
    mp.undo()
    print(f'max err {err:.4f} (CLS rows {err_cls:.4f}) OK')


if __name__ == '__main__':
    main()

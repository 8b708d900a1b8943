"""GPU (needs >= 2 devices, skipped otherwise): segment-sharded forward over NCCL gives bit-identical logits to one GPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, B, S, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from synchformer_b200 import model as M, ops, parallel, synth
    model = M.build_synchformer(n_segments=S, state_dict=synth.synthetic_state_dict(1337, n_segments=S), device=dev)
    vis = synth.synthetic_video(B, S, 0).half().view(B * S, 16, 3, 224, 224)
    wave = synth.synthetic_waveform(B, S, 0).view(B * S, 10240)
    s0, s1 = parallel.shard_range(B * S, world, rank)
    with torch.no_grad():
        mel = ops.mel_frontend(wave[s0:s1].contiguous().to(dev)).unsqueeze(1)
        logits = parallel.synchformer_forward_sharded(model, vis[s0:s1].contiguous().to(dev), mel, B, S)
        if rank == 0:
            full_mel = ops.mel_frontend(wave.to(dev)).view(B, S, 1, 128, 66)
            _, single = model(vis.view(B, S, 16, 3, 224, 224).to(dev), full_mel)
            q.put((logits.cpu(), single.cpu()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B,S', [(3, 2), (1, 3)])
def test_sharded_forward_is_bit_identical_to_single_gpu(B, S):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() + 13 * B + S) % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded, single = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sharded.shape == (B, 21)
    assert torch.equal(sharded, single)

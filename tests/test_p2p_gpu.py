"""The peer-memory all-gather (chr_p2p_allgather_f64, distributed.PeerGather) against NCCL's all_gather: needs two GPUs
(ranks that wait for one another must not share a device) - skipped on a one-GPU box."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=size, device_id=torch.device("cuda", rank))
    try:
        from chrysalis_b200 import distributed as dd
        peer = dd.PeerGather.create(4096)
        assert peer is not None, "symmetric memory is not available on this box"
        for call in range(7):                                          # both receive buffers, several epochs, ragged counts
            count = 4096 - 37 * call
            shard = torch.arange(count, dtype=torch.float64, device="cuda") * (rank + 1) + call
            got = peer.gather(shard, count)
            ref = [torch.empty(count, dtype=torch.float64, device="cuda") for _ in range(size)]
            dist.all_gather(ref, shard)
            for p in range(size):
                assert torch.equal(got[p, :count], ref[p]), (call, p)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_allgather_equals_nccl():
    size = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(size, _free_port()), nprocs=size, join=True)

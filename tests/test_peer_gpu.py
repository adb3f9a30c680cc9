"""Peer-memory collectives (csrc/peer_comm.cu) on one GPU: with world = 1 the kernels still run every phase
(stage into the slot, publish the call number to the own flag, poll, pull, scatter), so layout, alignment
fall-backs, slot alternation and CUDA-graph replay of the device-side call counter are checked here; the
multi-GPU agreement with NCCL is checked by `PeerComm.create`'s self-test and `tools/dp_check.py`."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def peer():
    from clear_vae_b200.peer import PeerComm
    pc = PeerComm(None, 0, 1, torch.device("cuda", 0), nbytes=4 << 20)
    yield pc
    pc.close()


def test_gather_is_identity_for_every_piece_kind(peer):
    g = torch.Generator().manual_seed(0)
    pieces = [torch.randn(1024, 8, generator=g).cuda(), torch.randn(1024, 2, generator=g).cuda(),
              torch.randint(-(1 << 40), 1 << 40, (1024,), generator=g).cuda(),
              torch.randn(333, 3, generator=g).cuda(),          # 3996 bytes: 4-byte path
              torch.randn(17, 8, generator=g).cuda()[1:]]        # 16-byte-multiple size at an unaligned-by-32 address
    for _ in range(3):   # both slots
        out = peer.gather(pieces)
        for a, b in zip(pieces, out):
            assert b.shape == a.shape and torch.equal(a, b)
    assert peer.error() == 0


def test_allreduce_world1_returns_the_inputs_and_replays_in_a_graph(peer):
    g = torch.Generator().manual_seed(1)
    ts = [torch.randn(n, generator=g).cuda() for n in (864, 32, 7, 18432, 1, 130, 5)]
    ts.append(torch.randn(66, generator=g).cuda()[1:])           # misaligned base: scalar scatter path
    ref = [t.clone() for t in ts]
    peer.allreduce_(ts)
    for a, b in zip(ts, ref):
        assert torch.equal(a, b)
    many = [torch.randn(3 + i, generator=g).cuda() for i in range(70)]   # > 64 tensors: split into two launches
    ref = [t.clone() for t in many]
    peer.allreduce_(many)
    for a, b in zip(many, ref):
        assert torch.equal(a, b)
    x = torch.randn(4096, 8, generator=g).cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        peer.gather([x])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        (y,) = peer.gather([x])
        peer.allreduce_([y])
    for i in range(5):
        x.add_(1.0)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(x, y)
    assert peer.error() == 0


def test_slot_overflow_is_reported(peer):
    big = torch.zeros(peer.slot_bytes() // 4 + 64, device="cuda")
    with pytest.raises(RuntimeError):
        peer.gather([big])

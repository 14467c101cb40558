"""CPU tier: the N > 1 path (circuits sharded round-robin over ranks, one all-gather of the readout rows) with
world_size = 2 over gloo. Each rank evolves its shard through the host API (CPU model primitives) and the
gathered table must equal the single-process result."""
import math
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
N_CIRCUITS, NQ = 5, 3


def _paths():
    for p in (ROOT, os.path.join(ROOT, 'tomography-assisted-mpdo-qcircuit_b200'), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


def readout_row(circuit_id):
    _paths()
    import MPDOSimulator as Simulator
    from MPDOSimulator import _engine, dmOperations
    from cpu_prims import CpuPrims
    _engine._PRIMS = CpuPrims()   # CPU model of the device primitives, injected by the test
    g = torch.Generator().manual_seed(1234 + circuit_id)
    c = Simulator.TensorCircuit(qn=NQ, ideal=False, noiseType='idealNoise', chi=4, kappa=2, chip='medium',
                                dtype=torch.complex128, device='cpu')
    for d in range(2):
        for q in range(NQ):
            th, ph, la = (torch.rand(3, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()
            c.u3(th, ph, la, [q])
        for q in range(d % 2, NQ - 1, 2):
            c.cz(q, q + 1)
    c.truncate()
    st = Simulator.Tools.create_ket0Series(NQ, dtype=torch.complex128)
    c.evolve(st)
    dmn = c.cal_dmNodes()
    row = [dmOperations.pauli_expect(dmn, 2, q).item() for q in range(NQ)]
    row.append(c.bitstring_probabilities(['0' * NQ])[0].item())
    return torch.tensor(row, dtype=torch.float64)


def _worker(rank, world, port, out_file):
    _paths()
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from MPDOSimulator._engine.sharding import gather_readout, shard
    mine = shard(range(N_CIRCUITS), rank, world)
    local = torch.stack([readout_row(i) for i in mine]) if mine else torch.zeros((0, NQ + 1), dtype=torch.float64)
    table = gather_readout(local, N_CIRCUITS)
    if rank == 0:
        torch.save(table, out_file)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    out_file = str(tmp_path / 'table.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out_file), nprocs=2, join=True)
    table = torch.load(out_file)
    want = torch.stack([readout_row(i) for i in range(N_CIRCUITS)])
    assert table.shape == want.shape
    assert (table - want).abs().max().item() < 1e-12

"""N>1 host logic on CPU: band partition, halo exchange and the rank-0 boundary-graph solves of
descriptools_b200/bands.py over torch.distributed (gloo, world_size 2), with the per-band device kernels
replaced by the NumPy stand-in tests/band_sim.py.  The solved seam values are checked against the oracle
run on the whole raster."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import oracle
from band_sim import KIND_FAIL, KIND_RIVER, flowacc_summary, hand_summary

PX, THR = 12.5, 60


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _weave(rows, cols, seam):
    r = np.arange(rows, dtype=np.float64)[:, None]
    c = np.arange(cols, dtype=np.float64)[None, :]
    centre = seam + 9.0 * np.sin(2 * np.pi * c / 40.0)
    z = 50.0 + 0.35 * np.abs(r - centre) + 0.05 * (cols - c) + 0.001 * ((r * 7 + c * 13) % 11)
    return oracle.priority_flood_eps(z.astype(np.float32))


def _case(kind):
    if kind == "weave":
        dem = _weave(128, 96, 64)
    else:
        dem = oracle.conditioned_dem(128, 80, seed=5)
        dem[20:30, 10:25] = -100
        dem = oracle.priority_flood_eps(dem)
    _, d8 = oracle.slope_d8(dem, PX)
    acc, left = oracle.flow_accumulation(d8)
    assert left == 0
    river = (acc > THR).astype(np.int8)
    fdist, idx, _ = oracle.flow_hand_index(dem, d8, river, PX)
    return dem, d8, acc, river, fdist, idx


def _worker(rank, world, port, kind, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch.distributed as dist

        dist.init_process_group("gloo", rank=rank, world_size=world)
        from descriptools_b200 import bands

        dem, d8, acc, river, fdist, idx = _case(kind)
        rows, cols = d8.shape
        edges = bands.band_edges(rows, world)
        r0, r1 = edges[rank], edges[rank + 1]
        x = bands.DistExchange()
        # D8 halo exchange
        mine = torch.from_numpy(d8[r0:r1].copy())
        above, below = torch.zeros(cols, dtype=torch.uint8), torch.zeros(cols, dtype=torch.uint8)
        x.halo([(mine[0].clone(), mine[-1].clone(), above, below)])
        if rank > 0:
            np.testing.assert_array_equal(above.numpy(), d8[r0 - 1])
        if rank + 1 < world:
            np.testing.assert_array_equal(below.numpy(), d8[r1])
        ha = above.numpy() if rank > 0 else None
        hb = below.numpy() if rank + 1 < world else None
        # flow accumulation: summaries -> rank-0 solve -> inflow of my halo rows
        fs = torch.from_numpy(flowacc_summary(d8[r0:r1], ha, hb))
        (inflow,) = x.solve([fs], bands.solve_flowacc_boundary)
        inflow = inflow.numpy()
        off = {1: (0, 1), 2: (1, 1), 4: (1, 0), 8: (1, -1), 16: (0, -1), 32: (-1, -1), 64: (-1, 0), 128: (-1, 1)}
        for side, hr, want_dr in ((0, r0 - 1, 1), (1, r1, -1)):
            if hr < 0 or hr >= rows:
                assert not inflow[side].any()
                continue
            for c in range(cols):
                code = int(d8[hr, c])
                into = code in off and off[code][0] == want_dr and 0 <= c + off[code][1] < cols and d8[hr + want_dr, c + off[code][1]] != 0
                if into:  # halo cell drains into my band: the solved inflow is its true acc + 1
                    assert inflow[side, c] == acc[hr, c] + 1, (rank, side, c)
        # HAND: summaries -> rank-0 pointer jumping -> resolved paths of my halo rows
        hs = torch.from_numpy(hand_summary(d8[r0:r1], ha, hb, river[r0:r1], dem[r0:r1], acc[r0:r1], r0))
        (res,) = x.solve([hs], bands.solve_hand_boundary)
        res = res.numpy()
        checked = 0
        for side, hr, want_dr in ((0, r0 - 1, -1), (1, r1, 1)):
            if hr < 0 or hr >= rows:
                continue
            brow = r0 if side == 0 else r1 - 1
            for c in range(cols):  # band cells that step onto halo cell (hr, c2)
                code = int(d8[brow, c])
                if code not in off or off[code][0] != want_dr:
                    continue
                c2 = c + off[code][1]
                if not (0 <= c2 < cols) or d8[hr, c2] == 0:
                    continue
                st = int(res[4 * side, c2])
                k, nd, nc = (st >> 62) & 3, (st >> 47) & 0x7FFF, (st >> 32) & 0x7FFF
                if idx[hr, c2] == -100:
                    assert k == KIND_FAIL
                else:
                    assert k == KIND_RIVER and res[4 * side + 1, c2] == idx[hr, c2]
                    np.testing.assert_allclose(nc * PX + nd * PX * np.sqrt(2.0), fdist[hr, c2], rtol=1e-5)
                    zr = np.array(res[4 * side + 2, c2]).view(np.float64)
                    assert zr == dem.flat[idx[hr, c2]] and res[4 * side + 3, c2] == acc.flat[idx[hr, c2]]
                checked += 1
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", checked))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "fail", traceback.format_exc()))


@pytest.mark.parametrize("kind", ["synth", "weave"])
def test_band_exchange_and_boundary_solves_gloo_world2(kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"
    assert sum(info for _, _, info in results) > 0  # some seam cells were really checked


def test_band_edges_on_tile_seams():
    from descriptools_b200 import bands

    assert bands.band_edges(40000, 8) == [0, 4992, 9984, 15040, 20032, 24960, 30016, 35008, 40000] or all(
        e % 64 == 0 for e in bands.band_edges(40000, 8)[1:-1])
    assert bands.band_edges(128, 2) == [0, 64, 128]
    with pytest.raises(ValueError):
        bands.band_edges(100, 4)


def test_solvers_three_bands_match_oracle_serial():
    """the same check without processes, three bands (LocalExchange): a path may cross two seams"""
    from descriptools_b200 import bands

    dem = _weave(192, 120, 64)
    _, d8 = oracle.slope_d8(dem, PX)
    acc, _ = oracle.flow_accumulation(d8)
    edges = [0, 64, 128, 192]
    summ = []
    for i in range(3):
        a, b = edges[i], edges[i + 1]
        summ.append(torch.from_numpy(flowacc_summary(d8[a:b], d8[a - 1] if i > 0 else None, d8[b] if i < 2 else None)))
    inflow = bands.LocalExchange(3).solve(summ, bands.solve_flowacc_boundary)
    assert not bool(bands.solve_flowacc_boundary(torch.stack(summ, 0))[1])  # resolved within the default rounds
    off = {1: (0, 1), 2: (1, 1), 4: (1, 0), 8: (1, -1), 16: (0, -1), 32: (-1, -1), 64: (-1, 0), 128: (-1, 1)}
    n = 0
    for i in range(3):
        for side, hr, want_dr in ((0, edges[i] - 1, 1), (1, edges[i + 1], -1)):
            if hr < 0 or hr >= 192:
                continue
            for c in range(120):
                code = int(d8[hr, c])
                if code in off and off[code][0] == want_dr and 0 <= c + off[code][1] < 120 and d8[hr + want_dr, c + off[code][1]] != 0:
                    assert int(inflow[i][side, c]) == acc[hr, c] + 1
                    n += 1
    assert n > 20

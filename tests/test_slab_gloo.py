"""The multi-GPU path on CPU: world_size-2 (and 3) gloo process groups run the y-slab driver
(voroffset_b200/slab.py: halo exchange, slab concatenation, row cropping) with a checker-backed backend
injected in place of the CUDA one. The product backend (CudaSlabBackend) is exercised on GPUs by
tests/test_gpu_parity.py::test_slab_single_process and bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleSlabBackend:
    """Test-only backend: host CompressedVolumes, compute by the oracle (never shipped)."""

    def __init__(self):
        from oracle.cpu import Oracle
        self.oracle = Oracle(threads=1)

    def shape(self, vol):
        return vol.nx, vol.ny

    def _rows(self, vol, y0, y1):
        c0, c1 = y0 * vol.nx, y1 * vol.nx
        off = vol.off[c0:c1 + 1].astype(np.int64)
        return (off - off[0]), vol.spans[off[0]:off[-1]]

    def new_tensors(self, n_off, n_spans):
        return torch.empty(n_off, dtype=torch.int32), torch.empty(max(2 * n_spans, 2), dtype=torch.float64)

    def rows_into(self, vol, y0, y1, off_t, spans_t):
        off, sp = self._rows(vol, y0, y1)
        off_t[: off.size] = torch.from_numpy(off.astype(np.int32))
        n = sp.shape[0]
        if n and 2 * n <= spans_t.numel():
            spans_t[: 2 * n] = torch.from_numpy(np.ascontiguousarray(sp).reshape(-1).copy())
        return n

    def from_tensors(self, nx, ny, off, spans, n_spans, like):
        return like.like(nx, ny, off.numpy()[: nx * ny + 1].astype(np.uint32), spans.numpy()[:2 * n_spans].reshape(-1, 2).copy())

    def concat(self, parts):
        parts = [p for p in parts if p is not None]
        offs, spans, base = [np.zeros(1, np.int64)], [], 0
        for p in parts:
            offs.append(p.off[1:].astype(np.int64) + base)
            base += int(p.off[-1])
            spans.append(p.spans)
        return parts[0].like(parts[0].nx, sum(p.ny for p in parts), np.concatenate(offs), np.concatenate(spans))

    def dilate_rows(self, vol, radius, y0, y1):
        full = self.oracle.morph3d(vol, "dilation", radius, "ours")
        off, sp = self._rows(full, y0, y1)
        return vol.like(vol.nx, y1 - y0, off, sp), 0.0, 0.0

    def release(self, vol):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, radius, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from voroffset_b200 import slab, synth
    vol = synth.blobs(44, count=30, padding=3, seed=13)          # same volume on every rank, each keeps its rows
    mine = slab.shard_rows(vol, rank, world)
    sd = slab.SlabDilation(OracleSlabBackend(), rank, world)
    out = sd.dilate(mine, radius)
    msgs = [sd.last_messages]
    # steady state: same link objects, one message batch per step; then a much denser volume (the halo
    # outgrows the agreed capacity -> overflow follow-up), then back
    out2 = sd.dilate(mine, radius)
    msgs.append(sd.last_messages)
    dense = slab.shard_rows(synth.random_volume(vol.nx, vol.ny, kmax=40, seed=5, zrange=400.0), rank, world)
    out3 = sd.dilate(dense, radius)
    msgs.append(sd.last_messages)
    out4 = sd.dilate(mine, radius)
    msgs.append(sd.last_messages)
    assert out2.bit_equal(out) and out4.bit_equal(out)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), off=out.off, spans=out.spans, ny=out.ny, halo=sd.last_halo_bytes,
             msgs=np.array(msgs), d_off=out3.off, d_spans=out3.spans)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,radius", [(2, 5.5), (2, 0.8), (3, 4.0)])
def test_slabs_reproduce_the_single_process_result(tmp_path, world, radius):
    from oracle.cpu import Oracle
    from voroffset_b200 import slab, synth
    mp.spawn(_worker, args=(world, _free_port(), radius, str(tmp_path)), nprocs=world, join=True)
    vol = synth.blobs(44, count=30, padding=3, seed=13)
    want = Oracle(threads=2).morph3d(vol, "dilation", radius, "ours")
    bounds = slab.slab_bounds(vol.ny, world)
    assert bounds[0][0] == 0 and bounds[-1][1] == vol.ny
    for r, (y0, y1) in enumerate(bounds):
        z = np.load(tmp_path / f"r{r}.npz")
        assert int(z["ny"]) == y1 - y0
        piece = slab.shard_rows(want, r, world)
        assert np.array_equal(z["off"], piece.off), f"rank {r}: topology differs"
        assert np.array_equal(z["spans"].view(np.uint64), piece.spans.view(np.uint64)), f"rank {r}: endpoints differ"
        if radius >= 1 and world > 1:
            assert int(z["halo"]) > 0
            msgs = z["msgs"].tolist()
            assert msgs[0] == 2 and msgs[1] == 1 and msgs[3] == 1      # count swap + payload, then one batch per step
            assert msgs[2] in (1, 2)                                    # overflow follow-up when the dense halo did not fit
        dense = synth.random_volume(vol.nx, vol.ny, kmax=40, seed=5, zrange=400.0)
        dpiece = slab.shard_rows(Oracle(threads=2).morph3d(dense, "dilation", radius, "ours"), r, world)
        assert np.array_equal(z["d_off"], dpiece.off) and np.array_equal(z["d_spans"].view(np.uint64), dpiece.spans.view(np.uint64))


def test_slab_bounds_cover_everything():
    from voroffset_b200 import slab
    for ny, w in [(10, 3), (2048, 8), (7, 7), (5, 1)]:
        b = slab.slab_bounds(ny, w)
        assert b[0][0] == 0 and b[-1][1] == ny and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(y1 - y0 for y0, y1 in b) - min(y1 - y0 for y0, y1 in b) <= 1


def test_thin_slabs_are_rejected():
    from voroffset_b200 import slab, synth
    sd = slab.SlabDilation(OracleSlabBackend(), rank=0, world=2)
    vol = synth.random_volume(8, 3, kmax=2, padding=0, seed=1)
    with pytest.raises(ValueError):
        sd.dilate(vol, 5.0)

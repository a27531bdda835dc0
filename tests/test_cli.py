"""The cyTVMPI-style command line (cytvdn_b200/cli.py): flags of mpi.py:47-76, block I/O on .npy files."""
import os

import numpy as np
import pytest

from cytvdn_b200 import cli
from cytvdn_b200.sharded import ShardPlan


def test_parser_accepts_the_reference_flags():
    a = cli.build_parser().parse_args("-i in.npy -o out.npy -d 4 -f 1 -n 40 -L .03 .03 .015 .015 -m 1 1 .5 .5 -v 0".split())
    assert a.dimensions == [4] and a.fista == [True] and a.niterations == [40] and a.verbose is False
    assert a.lam == [.03, .03, .015, .015] and a.mu == [1, 1, .5, .5]
    assert os.path.isabs(a.input[0]) and os.path.isabs(a.output[0])
    a = cli.build_parser().parse_args("-i a.npy -o b.npy -d 4 -n 10 5 -L 1 1 1 1 -m 1 1 1 1".split())
    assert a.niterations == [10, 5] and a.fista == [False]          # two counts -> hybrid (FISTA first)
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args("-o b.npy -d 4 -n 1 -L 1 -m 1".split())


def test_block_io_roundtrip(tmp_path):
    """Every rank reads its overlapped block and writes its owned block: together they rebuild the file."""
    rng = np.random.default_rng(0)
    g = rng.integers(0, 500, (11, 7, 4, 6)).astype(np.uint16)            # counts on disk, cast to float32 on read
    src, dst = str(tmp_path / "in.npy"), str(tmp_path / "out.npy")
    np.save(src, g)
    data = cli.open_input(src)
    cli.create_output(dst, g.shape)
    for rank in range(3):
        p = ShardPlan(g.shape, 3, rank)
        blk = cli.read_block(data, p.read_global)
        assert blk.dtype == np.float32 and blk.flags["C_CONTIGUOUS"] and blk.shape == p.local_shape
        cli.write_block(dst, p.owned_global, blk[p.owned_local] + 1)
    assert np.array_equal(np.load(dst), g.astype(np.float32) + 1)
    with pytest.raises(SystemExit, match="unsupported input format"):
        cli.open_input(str(tmp_path / "x.dm4"))


@pytest.mark.gpu
def test_cli_single_gpu_matches_api(tmp_path):
    import cytvdn_b200 as tv
    rng = np.random.default_rng(4)
    g = rng.poisson(rng.uniform(20, 400, (6, 7, 8, 16))).astype(np.float32)
    src, dst = str(tmp_path / "in.npy"), str(tmp_path / "out.npy")
    np.save(src, g)
    mu = np.array([1, 1, .5, .5], np.float32)
    lam = mu / 32
    argv = ["-i", src, "-o", dst, "-d", "4", "-f", "1", "-n", "9", "-v", "0", "-L"] + [str(float(v)) for v in lam] + \
           ["-m"] + [str(float(v)) for v in mu]
    assert cli.main(argv) == 0
    ref = tv.denoise4D(g, mu, 9, True, lam=lam, quiet=True)[0]
    assert np.array_equal(np.load(dst), ref)
    # 3-D, hybrid iteration counts
    c = rng.poisson(rng.uniform(20, 400, (6, 7, 32))).astype(np.float32)
    np.save(src, c)
    mu3 = np.array([1, 1, .5], np.float32)
    argv = ["-i", src, "-o", dst, "-d", "3", "-n", "4", "3", "-v", "0", "-L"] + [str(float(v)) for v in mu3 / 16] + \
           ["-m"] + [str(float(v)) for v in mu3]
    assert cli.main(argv) == 0
    ref = tv.denoise3D(c, mu3, [4, 3], lam=mu3 / 16, quiet=True)[0]
    assert np.array_equal(np.load(dst), ref)
    # out-of-core from the command line: forced with a tiny device budget, same result
    big = rng.poisson(rng.uniform(20, 400, (40, 5, 6, 16))).astype(np.float32)
    np.save(src, big)
    argv = ["-i", src, "-o", dst, "-d", "4", "-f", "1", "-n", "11", "-v", "0", "--schedule", "streamed", "-L"] + \
           [str(float(v)) for v in lam] + ["-m"] + [str(float(v)) for v in mu]
    os.environ["CYTVDN_STREAM_BUDGET_MB"] = repr(2.5 * 10 * 5 * 6 * 16 * 4 * 12 / 1048576.0 + 0.01)    # 12 planes per slot
    try:
        assert cli.main(argv) == 0
    finally:
        del os.environ["CYTVDN_STREAM_BUDGET_MB"]
    ref = tv.denoise4D(big, mu, 11, True, lam=lam, quiet=True)[0]
    assert np.array_equal(np.load(dst), ref)


class _FakeH5:
    """Just enough of h5py (File / Group / Dataset / attrs) to record what the EMD writer creates.  h5py itself is
    not installed in this image; the test pins the layout of `mpi.py:440-498`, not HDF5."""
    STORE = {}

    class _Attrs(dict):
        def create(self, k, v):
            self[k] = v

    class _Node:
        def __init__(self, store, path):
            self.store, self.path = store, path
            self.attrs = store.setdefault(("attrs", path), _FakeH5._Attrs())

        def create_group(self, name):
            p = f"{self.path}/{name}".lstrip("/")
            self.store[("group", p)] = True
            return _FakeH5._Node(self.store, p)

        def create_dataset(self, name, shape, dtype="float32"):
            p = f"{self.path}/{name}".lstrip("/")
            self.store[("data", p)] = np.zeros(shape, dtype=dtype)
            return _FakeH5._Dset(self.store, p)

        def __getitem__(self, name):
            p = f"{self.path}/{name}".lstrip("/")
            return _FakeH5._Dset(self.store, p) if ("data", p) in self.store else _FakeH5._Node(self.store, p)

    class _Dset(_Node):
        def __setitem__(self, sl, v):
            self.store[("data", self.path)][sl] = v

        @property
        def shape(self):
            return self.store[("data", self.path)].shape

    class File(_Node):
        def __init__(self, path, mode="r", **kw):
            if mode == "w":
                _FakeH5.STORE[path] = {}
            super().__init__(_FakeH5.STORE[path], "")

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False


def test_emd_writer_layout_with_fake_h5py(tmp_path, monkeypatch):
    """`-o x.emd`: the EMD v0.7 group structure the reference hard-codes (`mpi.py:440-490`) and block-wise writes of
    the owned tiles (`:493-497`), exercised through a stand-in for h5py."""
    import sys
    monkeypatch.setitem(sys.modules, "h5py", _FakeH5)
    dst = str(tmp_path / "out.emd")
    g = np.arange(11 * 7 * 4 * 6, dtype=np.float32).reshape(11, 7, 4, 6)
    assert cli.create_output(dst, g.shape) is None
    for rank in range(3):
        p = ShardPlan(g.shape, 3, rank)
        cli.write_block(dst, p.owned_global, g[p.owned_global])
    st = _FakeH5.STORE[dst]
    top = st[("attrs", "4DSTEM_experiment")]
    assert (top["emd_group_type"], top["version_major"], top["version_minor"]) == (2, 0, 7)
    for grp in ("metadata", "data", "data/datacubes", "data/counted_datacubes", "data/diffractionslices", "data/realslices",
                "data/pointlists", "data/pointlistarrays", "data/datacubes/datacube_0"):
        assert ("group", "4DSTEM_experiment/" + grp) in st, grp
    dc = "4DSTEM_experiment/data/datacubes/datacube_0"
    assert st[("attrs", dc)]["emd_group_type"] == 1 and st[("attrs", dc)]["metadata"] == -1
    assert np.array_equal(st[("data", dc + "/data")], g) and st[("data", dc + "/data")].dtype == np.float32
    for k, (n, label) in enumerate(zip(g.shape, (b"R_x", b"R_y", b"Q_x", b"Q_y")), start=1):
        assert np.array_equal(st[("data", f"{dc}/dim{k}")], np.arange(n))
        assert st[("attrs", f"{dc}/dim{k}")]["name"] == label and st[("attrs", f"{dc}/dim{k}")]["units"] == b"[pix]"
    with pytest.raises(SystemExit, match="4-D"):
        cli.create_output(dst, (4, 5, 6))
    monkeypatch.setitem(sys.modules, "h5py", None)           # import h5py -> ImportError
    with pytest.raises(SystemExit, match="needs h5py"):
        cli.create_output(dst, g.shape)

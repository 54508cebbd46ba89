"""GPU: the frequency-sharded map writer fed from device buffers, and the sharded makesky CLI (one rank)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_write_map_streams_device_blocks(tmp_path):
    import torch
    from cora_b200 import mapio

    freq = np.linspace(800.0, 400.0, 6, endpoint=False)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    blk = torch.randn((6, 4, 12 * 8 * 8), dtype=torch.float64, device="cuda", generator=gen)
    # a tiny staging ring: several chunks per block
    mapio.write_map(str(tmp_path / "a"), blk, freq, chunk_bytes=2 * 4 * 768 * 8)
    got, hdr = mapio.read_map(str(tmp_path / "a"))
    np.testing.assert_array_equal(got, blk.cpu().numpy())
    un = blk[:, 0].contiguous()
    mapio.write_map(str(tmp_path / "b"), un, freq, include_pol=True, chunk_bytes=768 * 8)
    got, hdr = mapio.read_map(str(tmp_path / "b"))
    np.testing.assert_array_equal(got[:, 0], un.cpu().numpy())
    assert got.shape == (6, 4, 768) and not got[:, 1:].any()


def test_makesky_sharded_cli_single_rank(tmp_path):
    """`cora_b200.makesky 21cm --sharded`: the sharded engine + writer; same maps as the dense CLI path for the seed."""
    from cora_b200 import makesky, mapio

    out = str(tmp_path / "sky")
    args = ["21cm", "--nside", "8", "--freq", "800", "700", "5", "--pol", "none", "--seed", "3"]
    makesky.main(args + ["--sharded", out])
    got, hdr = mapio.read_map(out)
    assert got.shape == (5, 1, 768) and hdr["pol"] == ["I"]
    np.testing.assert_allclose(hdr["freq"]["centre"], np.linspace(800.0, 700.0, 5, endpoint=False))
    assert hdr["freq"]["width"] == [20.0] * 5
    # statistics: zero-mean maps with the channel variance of the model (sum_l (2l+1)/(4 pi) C_l to cosmic variance)
    assert abs(got.mean()) < 5 * got.std() / np.sqrt(got.size) * 30
    makesky.main(["gaussianfg", "--nside", "8", "--freq", "800", "700", "3", "--pol", "full", "--seed", "1", "--sharded", out + "p"])
    gp, hp = mapio.read_map(out + "p")
    assert gp.shape == (3, 4, 768) and hp["pol"] == ["I", "Q", "U", "V"] and np.abs(gp[:, 1]).max() > 0 and not gp[:, 3].any()

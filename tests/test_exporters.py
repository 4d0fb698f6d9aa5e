"""Output path: tone mapping, PPM and OpenEXR writers -- CPU only."""
import numpy as np

from gpu_raytracer_b200 import exporters as ex


def test_tonemap_matches_the_shader_formula():
    c = np.array([[-1.0, 0.0, 0.18], [1.0, 4.0, 1e6]], dtype=np.float32)
    got = ex.tonemap_aces(c)
    x = np.maximum(c.astype(np.float64), 0.0)
    want = np.clip((x * (2.51 * x + 0.03)) / (x * (2.43 * x + 0.59) + 0.14), 0.0, 1.0) ** (1.0 / 2.2)
    assert np.allclose(got, want, atol=2e-6) and got[0, 0] == 0.0 and got[1, 2] == 1.0


def test_ppm_is_flipped_and_truncated(tmp_path):
    img = np.zeros((3, 2, 4), dtype=np.float32)
    img[0, 0, :3] = (1.0, 0.5, 0.999 / 255.0)      # bottom-left pixel of the device frame
    img[2, 1, :3] = (2.0, -1.0, 0.25)              # top-right, out of range on two channels
    p = str(tmp_path / "a.ppm"); ex.save_ppm(p, img)
    raw = open(p, "rb").read()
    assert raw.startswith(b"P6\n 2\n 3\n 255\n") and len(raw) == len(b"P6\n 2\n 3\n 255\n") + 18
    back = ex.load_ppm(p)
    assert tuple(back[0, 0]) == (255, 127, 0) and tuple(back[2, 1]) == (255, 0, 63)
    body = np.frombuffer(raw[-18:], dtype=np.uint8).reshape(3, 2, 3)
    assert tuple(body[0, 1]) == (255, 0, 63)        # the file's first row is the image's top row


def test_exr_half_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    img = (rng.random((5, 7, 4)) * 8.0).astype(np.float32)
    p = str(tmp_path / "a.exr"); ex.save_exr(p, img)
    raw = open(p, "rb").read()
    assert raw[:4] == bytes([0x76, 0x2F, 0x31, 0x01])
    assert raw.index(b"B\0") < raw.index(b"G\0") < raw.index(b"R\0")        # channel list in B, G, R order like the reference's
    back = ex.load_exr(p)
    assert back.shape == (5, 7, 3)
    assert np.array_equal(back, img[..., :3].astype(np.float16).astype(np.float32))     # exactly the HALF rounding, nothing else
    q = str(tmp_path / "b.exr"); ex.save_exr(q, img, half=False)
    assert np.array_equal(ex.load_exr(q), img[..., :3])
    expected = len(raw) - 5 * (8 + 3 * 7 * 2)
    assert raw[expected - 8 * 5:expected] == b"".join(int(expected + y * (8 + 42)).to_bytes(8, "little") for y in range(5))

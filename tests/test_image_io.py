"""Image file I/O of the reference's render(params) (src/lib.rs:57-71, src/color.rs:26-29, 215-231):
the engine's own PNG / PNM codec against Pillow as the independent codec, both directions, and the file-to-file render
on the GPU against the in-memory render of the same pixels."""
import os

import numpy as np
import pytest

from tests.helpers import noise_u8

PIL = pytest.importorskip("PIL.Image")


def _host():
    from film_grain_b200 import host as H
    return H


def _smooth(w, h):
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) % 256)], axis=2).astype(np.uint8)


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P", "1"])
@pytest.mark.parametrize("content", ["noise", "smooth"])
def test_png_written_by_pillow_decodes_to_the_same_rgb(tmp_path, mode, content):
    """Pillow picks scanline filters adaptively (smooth content exercises Sub / Up / Average / Paeth); every colour
    type the decoder accepts must come out as Pillow's own convert('RGB') -- alpha dropped, palette and 1-bit expanded."""
    H = _host()
    w, h = 67, 41  # odd sizes: partial bytes in the 1-bit rows
    rgb = noise_u8(w, h, seed=9) if content == "noise" else _smooth(w, h)
    im = PIL.fromarray(rgb, "RGB")
    if mode == "RGBA":
        im = im.convert("RGBA")
        im.putalpha(PIL.fromarray(rgb[:, :, 0]))
    elif mode == "P":
        im = im.quantize(colors=200)
    elif mode != "RGB":
        im = im.convert(mode)
    path = str(tmp_path / f"in_{mode}.png")
    im.save(path, optimize=(content == "smooth"))
    want = np.asarray(im.convert("RGB") if mode not in ("RGBA", "LA") else PIL.fromarray(np.asarray(im)[..., : (3 if mode == "RGBA" else 1)].squeeze()).convert("RGB"))
    got = H.load_image(path)
    assert got.shape == (h, w, 3)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("bits", [2, 4])
def test_low_bit_depth_grey_and_palette_png(tmp_path, bits):
    H = _host()
    w, h = 37, 19
    rng = np.random.default_rng(bits)
    idx = rng.integers(0, 1 << bits, (h, w), dtype=np.uint8)
    pal = rng.integers(0, 256, ((1 << bits), 3), dtype=np.uint8)
    im = PIL.fromarray(idx, "P")
    im.putpalette(pal.flatten().tolist())
    path = str(tmp_path / "pal.png")
    im.save(path, bits=bits)
    assert np.array_equal(H.load_image(path), pal[idx])


@pytest.mark.parametrize("fmt,ext", [("png", "png"), (None, "png"), ("ppm", "ppm"), (".PNG", "dat"), (None, "")])
def test_saved_files_are_read_back_by_pillow_and_by_the_engine(tmp_path, fmt, ext):
    H = _host()
    rgb = noise_u8(53, 31, seed=4)
    path = str(tmp_path / ("sub/dir/out" + ("." + ext if ext else "")))
    os.makedirs(os.path.dirname(path))
    H.save_image(path, rgb, fmt)
    assert np.array_equal(np.asarray(PIL.open(path).convert("RGB")), rgb)
    assert np.array_equal(H.load_image(path), rgb)
    head = open(path, "rb").read(8)
    assert head.startswith(b"P6") if (fmt == "ppm" or (fmt is None and ext == "ppm")) else head == b"\x89PNG\r\n\x1a\n"


def test_pnm_grey_with_comment_and_errors(tmp_path):
    H = _host()
    g = noise_u8(9, 5, seed=1)[:, :, 0]
    p = tmp_path / "g.pgm"
    p.write_bytes(b"P5\n# a comment\n9 5\n255\n" + g.tobytes())
    assert np.array_equal(H.load_image(str(p)), np.repeat(g[:, :, None], 3, axis=2))
    for name, data, msg in [("trunc.ppm", b"P6\n4 4\n255\n" + b"\0" * 10, "truncated"),
                            ("deep.ppm", b"P6\n1 1\n65535\n\0\0\0\0\0\0", "maxval"),
                            ("x.jpg", b"\xff\xd8\xff\xe0" + b"\0" * 20, "unsupported image format"),
                            ("bad.png", b"\x89PNG\r\n\x1a\n" + b"\0" * 30, "PNG")]:
        f = tmp_path / name
        f.write_bytes(data)
        with pytest.raises(H.RenderError, match=msg):
            H.load_image(str(f))
    with pytest.raises(H.RenderError, match="cannot open"):
        H.load_image(str(tmp_path / "missing.png"))
    im16 = PIL.fromarray((np.arange(12, dtype=np.uint16) * 5000).reshape(3, 4))
    im16.save(str(tmp_path / "d16.png"))
    with pytest.raises(H.RenderError, match="16-bit"):
        H.load_image(str(tmp_path / "d16.png"))
    # a flipped bit inside a chunk is caught by the CRC
    good = tmp_path / "good.png"
    H.save_image(str(good), noise_u8(8, 8, seed=2))
    raw = bytearray(good.read_bytes())
    raw[len(raw) // 2] ^= 0x10
    (tmp_path / "flip.png").write_bytes(bytes(raw))
    with pytest.raises(H.RenderError, match="CRC|corrupt"):
        H.load_image(str(tmp_path / "flip.png"))
    for token, msg in [("jpeg", "not built into"), ("xyz", "unsupported or unknown image format 'xyz'")]:
        with pytest.raises(H.RenderError, match=msg):
            H.save_image(str(tmp_path / "o.bin"), noise_u8(4, 4, seed=2), token)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_render_file_equals_the_in_memory_render(tmp_path, fused):
    """render(params): PNG in -> ROI crop -> device render -> directory created -> PNG / PPM out; the pixels are those of
    render_with_input_image on the decoded, cropped array (which the parity tests pin to the oracle)."""
    H = _host()
    img = noise_u8(96, 80, seed=12)
    src = str(tmp_path / "in.png")
    PIL.fromarray(img, "RGB").save(src)
    params = H.ParamsBuilder(radius_mean=0.1, n_samples=16, zoom=1.5, color_mode=H.ColorMode.Rgb, algo=H.Algo.Pixel).build()
    roi = (8, 4, 72, 60)
    want, _ = H.render_with_input_image(img[4:60, 8:72], params, fused=fused)
    for name, fmt in (("a/b/out.png", None), ("out.ppm", None), ("out.bin", "png")):
        dst = str(tmp_path / name)
        info = H.render(params, src, dst, output_format=fmt, roi=roi, fused=fused)
        assert (info.input_width, info.input_height) == (64, 56)
        assert np.array_equal(H.load_image(dst), want)
        assert np.array_equal(np.asarray(PIL.open(dst).convert("RGB")), want)
    with pytest.raises(H.RenderError, match="ROI exceeds image bounds"):
        H.render(params, src, str(tmp_path / "x.png"), roi=(0, 0, 97, 10))
    with pytest.raises(H.ParamsError):
        H.render(params, src, str(tmp_path / "x.png"), roi=(5, 5, 5, 10))
    with pytest.raises(H.RenderError, match="unsupported or unknown image format"):
        H.render(params, src, str(tmp_path / "x.qqq"))

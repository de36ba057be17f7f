"""PNG ingestion (the reference's loadTexture, staircase_scene.h:103-118, which decodes with the vendored
stb_image.h): host/png_reader.cpp against streams written HERE by an independent encoder (zlib from the standard library + the
filter and interlace rules of the PNG specification), so every filter type, colour type, bit depth, block type and Adam7 pass
is exercised, and against PIL's encoder when PIL is present. CPU only."""
import os
import struct
import zlib

import numpy as np
import pytest


def chunk(kind, body):
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)


def paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def filter_rows(rows, bpp, types):
    """rows: list of bytes (unfiltered scanlines). Returns the filtered stream with one filter-type byte per row."""
    out = bytearray()
    prev = bytes(len(rows[0])) if rows else b""
    for y, cur in enumerate(rows):
        t = types[y % len(types)]
        line = bytearray(len(cur))
        for i in range(len(cur)):
            a = cur[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, paeth(a, b, c))[t] if t < 5 else 0  # (t >= 5: an invalid type byte for the rejection test)
            line[i] = (cur[i] - pred) & 0xFF
        out.append(t)
        out += line
        prev = cur
    return bytes(out)


def pack_rows(samples, depth):
    """samples: (h, w*channels) integer array of `depth`-bit samples -> list of packed scanlines."""
    rows = []
    for r in samples:
        if depth == 8:
            rows.append(bytes(np.asarray(r, np.uint8)))
        elif depth == 16:
            rows.append(np.asarray(r, ">u2").tobytes())
        else:
            bits = "".join(format(int(v), "0%db" % depth) for v in r)
            bits += "0" * (-len(bits) % 8)
            rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
    return rows


def encode_png(samples, w, h, depth, colour, palette=None, interlace=False, types=(0, 1, 2, 3, 4), compress=None, split_idat=1):
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[colour]
    bpp = max(1, depth * channels // 8)
    s = np.asarray(samples).reshape(h, w, channels)
    if not interlace:
        raw = filter_rows(pack_rows(s.reshape(h, w * channels), depth), bpp, types)
    else:
        raw = b""
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = s[y0::dy, x0::dx]
            if sub.shape[0] == 0 or sub.shape[1] == 0:
                continue
            raw += filter_rows(pack_rows(sub.reshape(sub.shape[0], -1), depth), bpp, types)
    z = (compress or (lambda b: zlib.compress(b, 6)))(raw)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, colour, 0, 0, 1 if interlace else 0))
    if palette is not None:
        out += chunk(b"PLTE", bytes(np.asarray(palette, np.uint8).reshape(-1)))
    out += chunk(b"tEXt", b"Comment\0ancillary chunks are skipped")
    step = max(1, (len(z) + split_idat - 1) // split_idat)
    for i in range(0, len(z), step):
        out += chunk(b"IDAT", z[i:i + step])
    return out + chunk(b"IEND", b"")


def expected_rgb(samples, w, h, depth, colour, palette=None):
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[colour]
    s = np.asarray(samples, np.int64).reshape(h, w, channels)
    if colour == 3:
        return np.asarray(palette, np.uint8).reshape(-1, 3)[s[:, :, 0]]
    if depth == 16:
        s8 = s >> 8                                    # stb_image reduces 16-bit samples to their high byte
    elif depth < 8:
        s8 = s * (255 // ((1 << depth) - 1))           # 1/2/4-bit grey scaled to 0..255
    else:
        s8 = s
    s8 = s8.astype(np.uint8)
    return np.repeat(s8[:, :, :1], 3, axis=2) if channels <= 2 else s8[:, :, :3]


def fixed_huffman(b):
    c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    return c.compress(b) + c.flush()


CASES = [(8, 2), (8, 6), (8, 0), (8, 4), (16, 2), (16, 6), (16, 0), (16, 4), (1, 0), (2, 0), (4, 0), (8, 3), (4, 3), (2, 3), (1, 3)]


@pytest.mark.parametrize("depth,colour", CASES)
@pytest.mark.parametrize("interlace", [False, True])
def test_every_colour_type_bit_depth_filter_and_adam7(crt, depth, colour, interlace):
    rng = np.random.default_rng(depth * 10 + colour)
    w, h = 37, 21  # odd sizes: partial bytes at low bit depths, uneven Adam7 passes
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[colour]
    palette = rng.integers(0, 256, (1 << min(depth, 8), 3)) if colour == 3 else None
    # smooth content + noise so that the filters have something to predict
    yy, xx = np.mgrid[0:h, 0:w]
    base = ((xx * 5 + yy * 3)[:, :, None] * (np.arange(channels) + 1)) + rng.integers(0, 4, (h, w, channels))
    samples = base % (1 << depth)
    data = encode_png(samples, w, h, depth, colour, palette, interlace)
    got = crt.decode_png(data)
    assert got is not None and got.shape == (h, w, 3)
    assert np.array_equal(got, expected_rgb(samples, w, h, depth, colour, palette))
    assert np.array_equal(crt.decode_png(data, flip=True), got[::-1])  # stbi_set_flip_vertically_on_load(true)


@pytest.mark.parametrize("name,compress", [("stored", lambda b: zlib.compress(b, 0)), ("fixed", fixed_huffman),
                                           ("dynamic-1", lambda b: zlib.compress(b, 1)), ("dynamic-9", lambda b: zlib.compress(b, 9))])
def test_every_deflate_block_type_and_split_idat(crt, name, compress):
    rng = np.random.default_rng(7)
    w, h = 256, 200
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx + yy) % 256, (xx * 2) % 256, rng.integers(0, 256, (h, w))], axis=2)
    img[50:150, 30:200] = (200, 100, 50)  # long matches: distances and lengths with extra bits
    data = encode_png(img, w, h, 8, 2, compress=compress, split_idat=5)
    assert np.array_equal(crt.decode_png(data), img.astype(np.uint8))


def test_malformed_streams_are_rejected(crt):
    img = np.zeros((4, 4, 3), np.uint8)
    good = encode_png(img, 4, 4, 8, 2)
    assert crt.decode_png(good) is not None
    assert crt.decode_png(b"not a png at all") is None
    assert crt.decode_png(good[:40]) is None                               # truncated
    assert crt.decode_png(good.replace(b"IDAT", b"IDAX")) is None          # no image data
    bad_filter = encode_png(img, 4, 4, 8, 2, types=(7,))
    assert crt.decode_png(bad_filter) is None
    # a stream that inflates to far more than the header's 4x4 pixels allow (a "zip bomb") is cut off, not expanded
    bomb = good[:good.index(b"IDAT") - 4] + chunk(b"IDAT", zlib.compress(bytes(64 << 20), 9)) + chunk(b"IEND", b"")
    assert crt.decode_png(bomb) is None
    broken = bytearray(good)
    i = good.index(b"IDAT") + 4
    broken[i] = 0x79                                                       # zlib header check bits
    assert crt.decode_png(bytes(broken)) is None


def test_against_pil_encoder(crt, tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:96, 0:128]
    img = np.stack([(xx * 2) % 256, (yy * 2) % 256, ((xx + yy) % 64) * 4], axis=2).astype(np.uint8)
    img[20:40, 20:100] = rng.integers(0, 256, (20, 80, 3))
    for mode, arr in (("RGB", img), ("RGBA", np.dstack([img, np.full((96, 128), 128, np.uint8)])), ("L", img[:, :, 0]), ("P", None)):
        path = str(tmp_path / (mode + ".png"))
        pil = Image.fromarray(img).convert("P", palette=Image.ADAPTIVE, colors=64) if mode == "P" else Image.fromarray(arr, mode)
        pil.save(path, optimize=(mode == "RGB"))
        want = np.asarray(Image.open(path).convert("RGB"))
        got = crt.decode_png(open(path, "rb").read())
        assert got is not None and np.array_equal(got, want), mode


def test_scene_textures_from_png_files(crt, tmp_path):
    """loadTexture's contract: rows flipped, every byte / 255.0f; load_scene's nine file names in its order."""
    names = ["WoodFloor", "Wallpaper", "Woodpanel", "Painting1", "Painting2", "Painting3", "WoodChair", "Fabric", "BrushedAluminium"]
    rng = np.random.default_rng(11)
    imgs = []
    for k, n in enumerate(names):
        w, h = 8 + k, 5 + 2 * k
        img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        imgs.append(img)
        open(tmp_path / (n + ".png"), "wb").write(encode_png(img, w, h, 8, 2))
    scene = crt.Scene.staircase(0.05, 8, 5)
    before = scene.texture(3)
    assert scene.load_texture_png(3, str(tmp_path / "missing.png")) != 0 and np.array_equal(scene.texture(3), before)
    assert scene.load_texture_dir(str(tmp_path)) == 9
    for k, img in enumerate(imgs):
        tex = scene.texture(k)
        assert tex.shape == img.shape and tex.dtype == np.float32
        assert np.array_equal(tex, img[::-1].astype(np.float32) / np.float32(255.0))
    ks = scene.ks
    assert ks.numTextures == 9 and ks.textures[2].width == 10 and ks.textures[2].height == 9

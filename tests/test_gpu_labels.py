"""GPU parity of the label reader (csrc/labels.cu; CityLoader.py:86-96, :113-132 of the reference): bit-exact int64 maps
against the reference-made golden, the oracle (PIL + numpy) and PIL itself, through PNG files on disk."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import diga_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_city_loader_labels_golden(golden):
    from diga_b200.util.labels import resize_remap_labels, pseudo_label_lut, trainid_lut
    g = golden("city_loader_labels")
    crop = tuple(int(v) for v in g["crop_size"])
    lab = resize_remap_labels(torch.from_numpy(g["ids"]).to(DEV), crop, trainid_lut())
    pl = resize_remap_labels(torch.from_numpy(g["pseudo"]).to(DEV), crop, pseudo_label_lut(19))
    assert lab.dtype == torch.int64 and tuple(lab.shape) == crop
    assert np.array_equal(lab.cpu().numpy(), g["label_copy"]) and np.array_equal(pl.cpu().numpy(), g["pseudo_label_copy"])


@pytest.mark.parametrize("size0,crop", [((1024, 2048), (512, 1024)), ((1024, 2048), (512, 896)), ((1052, 1914), (512, 896)),
                                        ((64, 96), None), ((33, 47), (70, 95)), ((5, 3), (1, 1))])
def test_read_labels_from_png_files(tmp_path, size0, crop):
    """Palette PNG written by diga's own pseudo-label writer -> read_pseudo_label; labelIds PNG -> read_label."""
    from diga_b200.pseudolabel import colorize_mask
    from diga_b200.util.labels import read_label, read_pseudo_label
    rng = np.random.default_rng(size0[0])
    pl = rng.integers(0, 19, size0, dtype=np.uint8)
    pl[rng.random(size0) < 0.1] = 255
    ids = rng.integers(0, 34, size0, dtype=np.uint8)
    colorize_mask(pl).save(tmp_path / "pl.png")
    Image.fromarray(ids).save(tmp_path / "ids.png")
    got_pl = read_pseudo_label(str(tmp_path / "pl.png"), crop, device=DEV)
    got_id = read_label(str(tmp_path / "ids.png"), crop, device=DEV)
    ref_id, ref_pl = O.city_loader_labels(Image.open(tmp_path / "ids.png"), Image.open(tmp_path / "pl.png"), crop)
    assert np.array_equal(got_pl.cpu().numpy(), ref_pl) and np.array_equal(got_id.cpu().numpy(), ref_id)


def test_resize_remap_batch_and_errors():
    from diga_b200.util.labels import resize_remap_labels, pseudo_label_lut
    g = torch.Generator().manual_seed(1)
    src = torch.randint(0, 256, (3, 40, 50), generator=g, dtype=torch.uint8)
    out = resize_remap_labels(src.to(DEV), (17, 23))
    for i in range(3):
        ref = np.asarray(Image.fromarray(src[i].numpy()).resize((23, 17), Image.NEAREST)).astype(np.int64)
        ref = np.where(ref < 19, ref, 255)
        assert np.array_equal(out[i].cpu().numpy(), ref)
    with pytest.raises(RuntimeError):
        resize_remap_labels(src, (17, 23))                                  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        resize_remap_labels(src.to(DEV).long(), (17, 23))
    with pytest.raises(ValueError):
        resize_remap_labels(src.to(DEV), (17, 23), pseudo_label_lut()[:100])


# ------------------------------------------------------------------------------------------------ f3 writer: GPU deflate
def _png_patterns():
    rng = np.random.default_rng(5)
    coarse = rng.integers(0, 19, (9, 13)).astype(np.uint8)
    coarse[rng.random(coarse.shape) < 0.15] = 255
    out = {
        "blocks_odd": np.kron(coarse, np.ones((16, 16), np.uint8))[None, :131, :197],
        "constant": np.full((2, 40, 600), 7, np.uint8),
        "noise": rng.integers(0, 256, (2, 33, 70)).astype(np.uint8),
        "high_values": rng.integers(140, 256, (1, 17, 64)).astype(np.uint8),
        "one_pixel": np.array([[[200]]], np.uint8),
        "column": rng.integers(0, 19, (1, 50, 1)).astype(np.uint8),
        "run_258_edges": np.concatenate([np.full((1, 3, n), 5, np.uint8) for n in (257, 258, 259, 260, 261, 262)], axis=2),
        "wide_row": np.kron(rng.integers(0, 19, (1, 3, 40)).astype(np.uint8), np.ones((1, 1, 100), np.uint8)),
        "many_rows": np.kron(rng.integers(0, 19, (1, 300, 2)).astype(np.uint8), np.ones((1, 5, 9), np.uint8)),      # H > 1024
    }
    return out


@pytest.mark.parametrize("name", ["blocks_odd", "constant", "noise", "high_values", "one_pixel", "column", "run_258_edges",
                                  "wide_row", "many_rows"])
def test_png_deflate_stream_bit_exact_vs_oracle(name):
    """csrc/png.cu against oracle/png_oracle.py: the zlib stream of every image equal byte for byte, and standard zlib
    inflates it to the Up-filtered scanlines."""
    import zlib
    from oracle import png_oracle as P
    from diga_b200.pseudolabel import png_deflate
    lab = _png_patterns()[name]
    payload, lengths = png_deflate(torch.from_numpy(lab).to(DEV))
    payload, lengths = payload.cpu().numpy(), lengths.cpu().numpy()
    for i in range(lab.shape[0]):
        got = payload[i, :lengths[i]].tobytes()
        assert zlib.decompress(got) == P.filtered_scanlines(lab[i]).tobytes(), f"{name}[{i}] does not inflate to the scanlines"
        assert got == P.deflate_stream(lab[i]), f"{name}[{i}] differs from the oracle stream"


def test_png_deflate_full_resolution_batch_and_errors():
    """BASELINE config 5 size (1024x2048), a batch of segmentation-like maps plus one noisy map: inflate == scanlines for
    every image, oracle-equal for the first; buffer reuse (a second call over the same output) stays correct."""
    import zlib
    from oracle import png_oracle as P
    from diga_b200 import synthetic as S
    from diga_b200.pseudolabel import png_deflate
    g = S.gen(21, DEV)
    lab = S.block_labels(3, 1024, 2048, g, 32).to(torch.uint8)
    noise = torch.randint(0, 19, (1024, 2048), generator=g, device=DEV, dtype=torch.uint8)
    lab[2] = torch.where(torch.rand((1024, 2048), generator=g, device=DEV) < 0.05, noise, lab[2])
    for _ in range(2):
        payload, lengths = png_deflate(lab)
    host, lens = lab.cpu().numpy(), lengths.cpu().numpy()
    for i in range(3):
        got = payload[i, :lens[i]].cpu().numpy().tobytes()
        assert zlib.decompress(got) == P.filtered_scanlines(host[i]).tobytes()
    assert payload[0, :lens[0]].cpu().numpy().tobytes() == P.deflate_stream(host[0])
    assert lens[0] < 64 * 1024 < lens[2]
    with pytest.raises(RuntimeError):
        png_deflate(lab.cpu())
    with pytest.raises(ValueError):
        png_deflate(lab.long())


@pytest.mark.parametrize("encoder,prefix,coalesce", [("gpu", 256 * 1024, 1), ("gpu", 64, 1), ("gpu", 256 * 1024, 5), ("pil", 0, 1)])
def test_pseudo_label_writer_files_decode_like_the_reference(tmp_path, encoder, prefix, coalesce):
    """PseudoLabelWriter with the GPU encoder (including the long-stream second copy, forced by a 64-byte prefix) and with
    Pillow: every file opens as a 'P' image with the Cityscapes palette and the label map as pixel indices — what
    CityLoader.py:86-95 reads back."""
    from PIL import Image
    from diga_b200 import synthetic as S
    from diga_b200.pseudolabel import CITYSCAPES_PALETTE, PseudoLabelWriter
    g = S.gen(8, DEV)
    batches = [S.block_labels(2, 96, 160, g, 16).to(torch.uint8) for _ in range(6)]
    with PseudoLabelWriter(str(tmp_path), workers=3, slots=2, encoder=encoder, prefix=max(prefix, 1), coalesce=coalesce) as wr:
        for k, lab in enumerate(batches):
            wr.submit(lab, [f"x/y/im_{k}_{j}.png" for j in range(2)])
    assert wr.written == 12 and len(os.listdir(tmp_path)) == 12
    for k, lab in enumerate(batches):
        for j in range(2):
            png = Image.open(os.path.join(tmp_path, f"im_{k}_{j}.png"))
            assert png.mode == "P" and png.getpalette() == CITYSCAPES_PALETTE
            assert np.array_equal(np.array(png), lab[j].cpu().numpy())


def test_png_deflate_random_shapes_vs_oracle():
    """Seeded sweep over shapes / run structures (see the CPU twin in test_host_cpu.py): GPU stream == oracle stream."""
    from oracle import png_oracle as P
    from diga_b200.pseudolabel import png_deflate
    rng = np.random.default_rng(77)
    for _ in range(40):
        n, h, w = int(rng.integers(1, 4)), int(rng.integers(1, 60)), int(rng.integers(1, 900))
        n_runs = int(rng.integers(1, 12))
        row = np.repeat(rng.integers(0, 256, n_runs), rng.integers(1, 300, n_runs))[:w]
        row = np.pad(row, (0, w - row.size), mode="edge").astype(np.uint8)
        lab = np.tile(row, (n, h, 1))
        flip = rng.random((n, h, w)) < rng.choice([0.0, 0.01, 0.2])
        lab[flip] = rng.integers(0, 256, int(flip.sum()))
        payload, lengths = png_deflate(torch.from_numpy(lab).to(DEV))
        payload, lengths = payload.cpu().numpy(), lengths.cpu().numpy()
        for i in range(n):
            assert payload[i, :lengths[i]].tobytes() == P.deflate_stream(lab[i]), (n, h, w, i)


@pytest.mark.parametrize("name", ["blocks_odd", "constant", "noise", "one_pixel", "run_258_edges", "many_rows"])
def test_png_idat_crc_on_gpu_equals_zlib(name):
    """png_crc (csrc/png.cu png_crc_kernel: 256 ranges per image combined with the zlib x^n mod P operator) against
    zlib.crc32 over the chunk type and the stream, for streams from a few bytes to tens of kilobytes."""
    import zlib
    from diga_b200.pseudolabel import png_crc, png_deflate
    lab = _png_patterns()[name]
    payload, lengths = png_deflate(torch.from_numpy(lab).to(DEV))
    crc = png_crc(payload, lengths).cpu().numpy()
    payload, lengths = payload.cpu().numpy(), lengths.cpu().numpy()
    for i in range(lab.shape[0]):
        assert int(crc[i]) == zlib.crc32(b"IDAT" + payload[i, :lengths[i]].tobytes()), f"{name}[{i}]"


def test_png_idat_crc_arbitrary_bytes_and_lengths():
    """The CRC kernel on random bytes for every length class (0, shorter than the 256 ranges, not a multiple of them, MiB-sized)."""
    import zlib
    from diga_b200.pseudolabel import png_crc
    g = torch.Generator(device=DEV).manual_seed(3)
    cap = (1 << 20) + 77
    lens = [0, 1, 3, 4, 5, 255, 256, 257, 1023, 4096, 65537, 300001, cap]
    payload = torch.randint(0, 256, (len(lens), cap), generator=g, device=DEV, dtype=torch.uint8)
    lengths = torch.tensor(lens, device=DEV, dtype=torch.int64)
    crc = png_crc(payload, lengths).cpu().numpy()
    host = payload.cpu().numpy()
    for i, m in enumerate(lens):
        assert int(crc[i]) == zlib.crc32(b"IDAT" + host[i, :m].tobytes()), m
    with pytest.raises(ValueError):
        png_crc(payload, lengths[:3])
    with pytest.raises(RuntimeError):
        png_crc(payload.cpu(), lengths.cpu())

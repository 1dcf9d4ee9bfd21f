"""GPU parity of the label reader (csrc/labels.cu; CityLoader.py:86-96, :113-132 of the reference): bit-exact int64 maps
against the reference-made golden, the oracle (PIL + numpy) and PIL itself, through PNG files on disk."""
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

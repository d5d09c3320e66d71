"""Golden vectors for the device-side batch synthesis: runs the UNMODIFIED reference
datasets/DenoisingDatasets.py:SimulateTrain.__getitem__ on CPU in the build container over three small PNG files
written to a temp directory, and stores the images plus every sample's outputs.  The per-sample draws are replayable:
`replay_draws` consumes Python's `random` and torch's generator in the reference's order.
   python tools/gen_golden_synth.py"""
import random
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

OUT = ROOT / "tests" / "golden"
PCH = 32
SEEDS = {"niid": [3, 4, 5, 6, 7, 8, 9, 10, 11, 12], "iid": [21, 22], "niid_clip": [31, 32, 33]}


def make_images():
    rng = np.random.default_rng(77)
    return [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for h, w in ((48, 56), (40, 33), (64, 64))]


def replay_draws(seed, images, mode, pch=PCH):
    """The reference's consumption of `random` / torch RNG for one __getitem__ after reset_seed(seed):
    image index, crop offsets, sigma-map scalars, torch.randn noise, augmentation flag."""
    random.seed(seed)
    torch.manual_seed(seed)
    ind_im = random.randint(0, len(images) - 1)
    im = images[ind_im]
    ind_h = random.randint(0, im.shape[0] - pch)
    ind_w = random.randint(0, im.shape[1] - pch)
    patch = im[ind_h:ind_h + pch, ind_w:ind_w + pch]
    if mode.startswith("niid"):
        center = [random.uniform(0, pch), random.uniform(0, pch)]
        scale = random.uniform(pch / 4, pch / 4 * 3)
        up = random.uniform(0 / 255.0, 75 / 255.0)
        down = random.uniform(0 / 255.0, 75 / 255.0)
        if up < down:
            up, down = down, up
        up += 5 / 255.0
        params = [center[0], center[1], scale, up, down, 0.0]
    else:
        params = [0.0, 0.0, -1.0, 0.0, 0.0, random.uniform(0 / 255.0, 75 / 255.0)]
    noise = torch.randn(patch.shape).numpy()
    aug = random.randint(0, 7)
    return patch, params, aug, noise


def main():
    import cv2
    import ref_import
    ref_import.import_reference()
    import importlib
    import types
    for missing in ("h5py", "lmdb"):                              # only the real-noise (h5 / lmdb) datasets need them
        sys.modules.setdefault(missing, types.ModuleType(missing))
    dd = importlib.import_module("datasets.DenoisingDatasets")
    # scikit-image is absent here; its uint8 -> float32 conversion is np.multiply(image, 1. / 255, dtype=float32)
    # (skimage/util/dtype.py:_convert), which differs from a division by one ulp for some values: mirror it exactly
    dd.img_as_float32 = lambda x: np.multiply(np.asarray(x), 1.0 / 255, dtype=np.float32)
    images = make_images()
    out = {"images": [torch.from_numpy(im.copy()) for im in images], "samples": {}}
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i, im in enumerate(images):
            pth = str(Path(tmp) / f"im{i}.png")
            cv2.imwrite(pth, im[:, :, ::-1])           # the dataset flips BGR -> RGB on read
            paths.append(pth)
        for mode, seeds in SEEDS.items():
            ds = dd.SimulateTrain(paths, length=16, pch_size=PCH, chn=3, mode=mode.split("_")[0], clip=mode.endswith("clip"))
            for seed in seeds:
                ds.reset_seed(seed)
                im_noisy, im_gt, sigma_gt = ds[0]
                out["samples"][(mode, seed)] = (im_noisy.clone(), im_gt.clone(), sigma_gt.clone())
    torch.save(out, OUT / "synth_denoise.pt")
    print("saved", len(out["samples"]), "samples")


if __name__ == "__main__":
    main()

"""Generates tests/golden/preprocess_cases.npz with the SAME library calls as the reference's `preprocess_for_sr`
(inference/run.py:80-96: cv2.cvtColor BGR2RGB -> PIL Image.resize((192, 32), Image.BICUBIC) -> torchvision ToTensor -> unsqueeze(0)).
inference/run.py itself cannot be imported headless (it pulls in the GUI / YOLO stack), so the four lines are invoked directly.
Run in the build container (cv2, Pillow, torchvision present): python tests/golden/make_golden_preprocess.py"""
import os

import cv2
import numpy as np
import PIL
import torchvision
import torchvision.transforms as T
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_preprocess_for_sr(plate_image, target_size=(192, 32)):
    rgb_img = cv2.cvtColor(plate_image, cv2.COLOR_BGR2RGB)
    pil_img = Image.fromarray(rgb_img).resize(target_size, Image.BICUBIC)
    transform = T.Compose([T.ToTensor()])
    return transform(pil_img).unsqueeze(0)


if __name__ == "__main__":
    rng = np.random.default_rng(2024)
    sizes = [(32, 192), (20, 90), (33, 191), (64, 300), (11, 47), (48, 120), (100, 400), (16, 192), (40, 33), (3, 5), (1, 1), (32, 100)]
    out = {"versions": np.array([cv2.__version__, PIL.__version__, torchvision.__version__])}
    for i, (h, w) in enumerate(sizes):
        if i % 3 == 2:   # smooth content (plates are not noise): low-frequency pattern + noise
            yy, xx = np.mgrid[0:h, 0:w]
            base = 127 + 100 * np.sin(xx / 7.0)[..., None] * np.cos(yy / 5.0)[..., None] * np.array([1.0, 0.6, -0.8])
            img = np.clip(base + rng.normal(0, 8, (h, w, 3)), 0, 255).astype(np.uint8)
        else:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out[f"in_{i}"] = img
        out[f"out_{i}"] = reference_preprocess_for_sr(img).numpy()
    np.savez_compressed(os.path.join(HERE, "preprocess_cases.npz"), **out)
    print("wrote", len(sizes), "cases")

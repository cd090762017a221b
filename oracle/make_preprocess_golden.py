"""Generates tests/golden/preprocess_digests.json: sha256 digests of what the reference's preprocessing dependencies
produce in this container -- PIL.Image (Pillow, bicubic resize of the expand2square'd frame) followed by the
transformers-4.44.2 rescale / normalize arithmetic in numpy -- for seeded synthetic frames.  Run from the repo root:
    python oracle/make_preprocess_golden.py
tests/test_preprocess_cpu.py checks oracle/preprocess.py against these digests (and against PIL itself when importable)."""
import hashlib, json, os, sys

import numpy as np
from PIL import Image
import PIL

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from preprocess_cases import CASES, make_frame      # noqa: E402

MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)


def expand2square(pil_img, background_color):              # the reference's helper, restated on PIL objects
    width, height = pil_img.size
    if width == height:
        return pil_img
    side = max(width, height)
    result = Image.new(pil_img.mode, (side, side), background_color)
    result.paste(pil_img, (0, (width - height) // 2) if width > height else ((height - width) // 2, 0))
    return result


out = {"pillow": PIL.__version__, "cases": {}}
for name, (h, w, seed) in CASES.items():
    img = expand2square(Image.fromarray(make_frame(h, w, seed)), tuple(int(x * 255) for x in MEAN))
    u8 = np.asarray(img.resize((336, 336), Image.BICUBIC))
    x = (u8 * (1 / 255)).astype(np.float32)
    x = ((x - np.array(MEAN, dtype=np.float32)) / np.array(STD, dtype=np.float32)).transpose(2, 0, 1)
    out["cases"][name] = {"u8_sha256": hashlib.sha256(np.ascontiguousarray(u8).tobytes()).hexdigest(),
                          "f16_sha256": hashlib.sha256(np.ascontiguousarray(x).astype(np.float16).tobytes()).hexdigest()}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "preprocess_digests.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)

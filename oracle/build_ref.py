"""TEST / MEASUREMENT INFRASTRUCTURE: build oracle/_ref -- a pruned, importable view of the UNMODIFIED reference.

The reference (/root/reference) is pure Python: there is nothing to compile, but it cannot travel to the GPU box, and
`bench.py --impl reference` / the `gpu_baseline` leg must time the reference's OWN `ddim_sample_loop` +
`ClassifierFreeSampleModel` + `FiLMTransformer.forward` there.  This recipe copies, byte for byte, the Python modules
that path imports (and nothing else: no renderer, no training code, no assets) into oracle/_ref/, which is git-ignored
(never part of this repository's history) but NOT gpurun-ignored, so it travels with the snapshot like a built .so.

    python -m oracle.build_ref          # in the build container (needs /root/reference)

oracle/ref_harness.py resolves the reference root as /root/reference when present, else oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# modules on the import closure of: diffusion.{gaussian_diffusion,respace}, model.{diffusion,cfg_sampler,guide,vqvae},
# utils.model_util, sample.generate (entry point; imported, never run as a CLI)
FILES = [
    "diffusion/__init__.py", "diffusion/gaussian_diffusion.py", "diffusion/respace.py", "diffusion/losses.py", "diffusion/nn.py",
    "model/__init__.py", "model/diffusion.py", "model/cfg_sampler.py", "model/guide.py", "model/vqvae.py", "model/utils.py",
    "model/modules/__init__.py", "model/modules/transformer_modules.py", "model/modules/rotary_embedding_torch.py",
    "model/modules/audio_encoder.py",
    "utils/__init__.py", "utils/model_util.py", "utils/misc.py", "utils/diff_parser_utils.py",
    "sample/__init__.py", "sample/generate.py",
    "data_loaders/__init__.py", "data_loaders/get_data.py", "data_loaders/data.py", "data_loaders/tensors.py",
    "LICENSE",
]


def build(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(SRC, "diffusion")):
        if os.path.isdir(os.path.join(DST, "diffusion")):
            return DST          # GPU box: use the prebuilt view
        raise SystemExit("oracle/build_ref.py needs /root/reference (build container only)")
    manifest = {}
    os.makedirs(DST, exist_ok=True)
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            if rel.endswith("__init__.py"):       # namespace packages in the reference: keep them namespace packages
                continue
            raise SystemExit(f"reference file missing: {rel}")
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1)
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""Stage the UNMODIFIED reference under oracle/_ref/ (git-ignored, NOT gpurun-ignored) so that it travels to the GPU box
with the repo snapshot, like the built .so.  Only the Python modules (+ the BPE vocabulary the CLIP tokenizer reads at
import) that the eval forward and the eval caller import are staged — byte-for-byte copies, never edited, never committed.

    python oracle/stage_ref.py            # needs /root/reference; idempotent

Users: bench.py's reference arm / cpu_baseline (`kind: "reference"`) and its same-GPU torch comparator, through
oracle/ref_harness.py (which resolves /root/reference first, then oracle/_ref).  Product code never imports either.
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"

FILES = [
    "upt_tip_cache_model_free_finetune_distill3.py", "CLIP_models_adapter_prior2.py", "ops.py", "hico_list.py",
    "hico_text_label.py", "vcoco_list.py",
    "CLIP/clip/__init__.py", "CLIP/clip/clip.py", "CLIP/clip/model.py", "CLIP/clip/simple_tokenizer.py",
    "CLIP/clip/bpe_simple_vocab_16e6.txt.gz",
    "detr/models/__init__.py", "detr/models/backbone.py", "detr/models/detr.py", "detr/models/matcher.py",
    "detr/models/position_encoding.py", "detr/models/segmentation.py", "detr/models/transformer.py",
    "detr/util/__init__.py", "detr/util/box_ops.py", "detr/util/misc.py",
    # the eval caller (f1 / f2 oracles): CustomisedDLE.test_hico's association + AP meter
    "pocket/pocket/utils/association.py", "pocket/pocket/utils/meters.py",
]


def stage(verbose: bool = True) -> bool:
    if not (SRC / FILES[0]).exists():
        if verbose:
            print(f"[stage_ref] {SRC} not present: nothing staged (oracle/_ref {'exists' if DST.exists() else 'absent'})")
        return DST.exists()
    for rel in FILES:
        src, dst = SRC / rel, DST / rel
        if not src.exists():
            raise FileNotFoundError(src)
        dst.parent.mkdir(parents=True, exist_ok=True)
        if not dst.exists() or dst.read_bytes() != src.read_bytes():
            shutil.copyfile(src, dst)
    if verbose:
        print(f"[stage_ref] staged {len(FILES)} reference files under {DST} ({sum((DST / f).stat().st_size for f in FILES) / 1e6:.1f} MB)")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)

"""Stage the reference's hot-path files (the starred rows) into oracle/_ref/ so that the GPU box
can run the UNMODIFIED reference as the CPU baseline.

`/root/reference` exists only in the build container; `gpurun` snapshots /root/repo, and `oracle/_ref/` is
git-ignored (never committed - the repo holds no reference source) but NOT gpurun-ignored, so the staged copy
travels exactly like the in-tree `libb2dglue.so`.  `oracle/ref_import.py` imports from `/root/reference/lib`
when it exists and from `oracle/_ref/lib` otherwise.  Test infrastructure only: nothing under
`faster_rcnn_pytorch_multimodal_b200/` may import it (tests/test_abi.py checks).

    python -m oracle.stage_ref        # called by __graft_entry__.build() when /root/reference is present
"""
import os
import shutil

SRC = "/root/reference/lib"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "lib")
FILES = [
    "__init__.py",
    "layer_utils/__init__.py", "layer_utils/proposal_layer.py", "layer_utils/proposal_top_layer.py",
    "layer_utils/anchor_target_layer.py", "layer_utils/proposal_target_layer.py", "layer_utils/generate_anchors.py",
    "layer_utils/generate_3d_anchors.py", "layer_utils/snippets.py",
    "model/__init__.py", "model/config.py", "model/bbox_transform.py",
    "utils/__init__.py", "utils/bbox.py", "utils/torchpoolers.py", "utils/loss_utils.py", "utils/filter_predictions.py",
]


def stage() -> str:
    if not os.path.isdir(SRC):
        raise RuntimeError(f"{SRC} is not present: staging only works in the build container")
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
        else:
            raise RuntimeError(f"reference file missing: {src}")
    return DST


if __name__ == "__main__":
    print(stage())

"""Recipe for ``oracle/_ref/``: a travelling copy of the LIVE reference package, so that the GPU box (where
``/root/reference`` does not exist) can time the reference's own ``Seq2SeqModel.inference`` /
``_greedy_search`` (model/seq2seq.py:124-232, with its per-row ``.item()`` host syncs) as the
``bench.py --impl reference`` arm and as ``cpu_baseline`` (kind "reference").

TEST / MEASUREMENT INFRASTRUCTURE -- see ``oracle/__init__.py``.  The reference is pure Python: there is
nothing to compile; the "build" is a verbatim copy of ``/root/reference/img2latex`` into ``oracle/_ref/img2latex``.
``oracle/_ref/`` is git-ignored (reference sources never enter the history) but not gpurun-ignored, so it rides along
with the snapshot like the built ``.so``.  Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/img2latex"
DST = os.path.join(HERE, "_ref", "img2latex")


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present (GPU box?): keeping whatever oracle/_ref holds")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "outputs", "*.png", "*.jpg"))
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    print(f"make_ref: copied {n} files of the reference package to {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

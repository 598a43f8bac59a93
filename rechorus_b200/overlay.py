"""Run the reference's UNCHANGED ``src/main.py`` with the kernel-backed model classes substituted.

    python -m rechorus_b200.overlay --ref /root/reference/src --model_name BPRMF --emb_size 64 --lr 1e-3 \
        --l2 1e-6 --dataset Grocery_and_Gourmet_Food --path /tmp/data/ [--table_mode fused]

What it does: puts the reference ``src/`` on ``sys.path``, restores the NumPy aliases the
reference still uses (np.object / np.int / np.float), imports the reference's model modules, builds classes
that graft this package's kernel mixins onto the reference's own ``GeneralModel`` / ``SequentialModel`` (so the
reference's Dataset, reader, runner, argument plumbing and checkpoint code are the ones that run), replaces
``models.general.BPRMF.BPRMF`` etc. with them, and ``runpy``s ``main.py``.  Needs a CUDA device for the first
forward call; there is no CPU fallback.
"""
from __future__ import annotations

import os
import runpy
import sys

DEFAULT_REF = "/root/reference/src"


def install(ref_src: str = DEFAULT_REF):
    import numpy as np
    for alias, typ in (("object", object), ("int", int), ("float", float)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)          # removed in NumPy >= 1.24; BaseModel.py:141, SASRec.py:69, utils.py:65
    if ref_src not in sys.path:
        sys.path.insert(0, ref_src)
    import models.BaseModel as RB                       # the reference's own base classes
    import models.general.BPRMF as ref_bprmf
    import models.general.NeuMF as ref_neumf
    import models.sequential.SASRec as ref_sasrec
    from . import plugin

    class BPRMF(plugin.BPRMFKernels, RB.GeneralModel):
        reader, runner = "BaseReader", "BaseRunner"
        extra_log_args = ["emb_size", "batch_size"]

        @staticmethod
        def parse_model_args(parser):
            parser = plugin.BPRMFKernels.parse_model_args(parser)
            return RB.GeneralModel.parse_model_args(parser)

        def __init__(self, args, corpus):
            RB.GeneralModel.__init__(self, args, corpus)
            self._base_init(args, corpus)

    class NeuMF(plugin.NeuMFKernels, RB.GeneralModel):
        reader, runner = "BaseReader", "BaseRunner"
        extra_log_args = ["emb_size", "layers"]

        @staticmethod
        def parse_model_args(parser):
            parser = plugin.NeuMFKernels.parse_model_args(parser)
            return RB.GeneralModel.parse_model_args(parser)

        def __init__(self, args, corpus):
            RB.GeneralModel.__init__(self, args, corpus)
            self._neumf_init(args, corpus)

    class SASRec(plugin.SASRecKernels, RB.SequentialModel):
        reader, runner = "SeqReader", "BaseRunner"
        extra_log_args = ["emb_size", "num_layers", "num_heads"]

        @staticmethod
        def parse_model_args(parser):
            parser = plugin.SASRecKernels.parse_model_args(parser)
            return RB.SequentialModel.parse_model_args(parser)

        def __init__(self, args, corpus):
            RB.SequentialModel.__init__(self, args, corpus)
            self._base_init(args, corpus)

    ref_bprmf.BPRMF, ref_neumf.NeuMF, ref_sasrec.SASRec = BPRMF, NeuMF, SASRec
    return {"BPRMF": BPRMF, "NeuMF": NeuMF, "SASRec": SASRec}


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    ref = DEFAULT_REF
    if "--ref" in argv:
        i = argv.index("--ref")
        ref = argv[i + 1]
        del argv[i:i + 2]
    install(ref)
    sys.argv = [os.path.join(ref, "main.py")] + argv
    runpy.run_path(os.path.join(ref, "main.py"), run_name="__main__")


if __name__ == "__main__":
    main()

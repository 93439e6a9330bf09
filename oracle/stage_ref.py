"""TEST INFRASTRUCTURE ONLY — stages the UNMODIFIED reference files of the hot path into oracle/_ref/.

The reference (G-U-N/consolver) is pure Python, so "building" it is a byte-for-byte copy of the few source
files on the path.  They go to the git-ignored `oracle/_ref/` (never into history), which — like the in-tree
`libconsolver.so` — travels to the GPU box with the working tree, where /root/reference does not exist.
`oracle/ref_shim.py` loads them from there (stand-ins only for the absent `diffusers` base classes), so that

  * `bench.py --impl reference` and `cpu_baseline` time `PPOScheduler.step` ITSELF on the box's host cores,
  * `bench.py`'s `torch_eager_gpu` / `torch_compile_gpu` legs run the same classes on cuda:0 (SURVEY §2.2's bar),
  * `oracle/make_golden.py --device cuda` writes goldens from the reference running on a real B200.

`MANIFEST.json` records the sha256 of every staged file next to the sha256 of its source, so a reader can check
that nothing was edited.  Run:  python oracle/stage_ref.py   (done by __graft_entry__.build() when the
reference tree is present).  Nothing under consolver_b200/ reads oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("CONSOLVER_REFERENCE_SRC", "/root/reference")

# the files ref_shim.load_reference() executes: the two schedulers, their policies, the baseline solvers — and the
# SD caller loop, which tests/test_gpu_live_reference.py runs unmodified over BOTH schedulers
FILES = [
    "scheduler_ppo.py",                    # PPOScheduler                     
    "factor_net_ppo.py",                   # FactorNetPPO, SD                 (F1-F4, T1)
    "conv_net.py",                         # imported by factor_net_ppo.py
    "edit_ppo/scheduler_fmppo.py",         # FMPPOScheduler                   (M0-M2)
    "edit_ppo/factor_net_ppo.py",          # FactorNetPPO, FM
    "edit_ppo/conv_net.py",
    "edit_ppo/scheduler_fm.py",            # FlowMatchGeneralDiscreteScheduler (N4 baselines)
    "diffusers_amed_plugin_dpmpp.py",      # AMED plugin                      (N4)
    "denoise_ppo.py",                      # the caller loop (D1, D2): drives the scheduler in the live drop-in test
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(source: str = SOURCE, dest: str = DEST, quiet: bool = False) -> bool:
    """Copy FILES from `source` to `dest`.  Returns False (and leaves `dest` alone) when `source` is absent."""
    if not os.path.isfile(os.path.join(source, FILES[0])):
        return False
    manifest = {"source": source, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(source, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        a, b = _sha(src), _sha(dst)
        if a != b:
            raise RuntimeError(f"staged copy of {rel} differs from its source")
        manifest["files"][rel] = a
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(FILES)} unmodified reference files into {dest}")
    return True


def verify(dest: str = DEST) -> bool:
    """True when every file listed in dest/MANIFEST.json is present with the recorded sha256."""
    mf = os.path.join(dest, "MANIFEST.json")
    if not os.path.isfile(mf):
        return False
    with open(mf) as f:
        files = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(dest, rel)) and _sha(os.path.join(dest, rel)) == h
               for rel, h in files.items())


if __name__ == "__main__":
    ok = stage()
    if not ok:
        print(f"reference tree not found at {SOURCE}; nothing staged", file=sys.stderr)
        sys.exit(0 if verify() else 1)

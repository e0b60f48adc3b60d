"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference's hot-path modules available where the
reference tree is not mounted (the GPU box): the files SURVEY.md section 8(a) cites, plus the modules they import from
their own package, are copied byte for byte from ``/root/reference`` into the git-ignored ``oracle/_ref/`` (which is
NOT gpurun-ignored, so it travels with the snapshot like the built ``.so``).  Nothing is edited; nothing under
``oracle/_ref/`` is ever committed (see ``.gitignore``).

Used by ``oracle/ref_shim.py`` as the fall-back reference root, so that

  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the real ``VFAgent.update`` (kind "reference"), and
  * ``-m gpu`` tests and ``smoke()`` can compare the CUDA path with the live reference, not only with the restatement.

``build()`` is called by ``__graft_entry__.build()``; without ``/root/reference`` it does nothing (the GPU box only uses
what was copied in the build container).
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ABM_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

# hot-path files and the in-package modules they import at module level
FILES = [
    ".env",
    "abm/__init__.py",
    "abm/agent/__init__.py", "abm/agent/agent.py", "abm/agent/supcalc.py",
    "abm/contrib/__init__.py", "abm/contrib/colors.py", "abm/contrib/decision_params.py",
    "abm/contrib/movement_params.py", "abm/contrib/ifdb_params.py", "abm/contrib/evolution.py",
    "abm/environment/__init__.py", "abm/environment/rescource.py",
    "abm/simulation/__init__.py", "abm/simulation/sims.py", "abm/simulation/interactions.py",
    "abm/monitoring/__init__.py", "abm/monitoring/ifdb.py", "abm/monitoring/env_saver.py",
    "abm/loader/__init__.py", "abm/loader/helper.py", "abm/loader/data_loader.py",
    "abm/projects/__init__.py",
    "abm/projects/visual_flocking/__init__.py",
    "abm/projects/visual_flocking/vf_agent/__init__.py",
    "abm/projects/visual_flocking/vf_agent/vf_agent.py",
    "abm/projects/visual_flocking/vf_agent/vf_supcalc.py",
    "abm/projects/visual_flocking/vf_contrib/__init__.py",
    "abm/projects/visual_flocking/vf_contrib/vf_params.py",
    "abm/projects/visual_flocking/vf_simulation/__init__.py",
    "abm/projects/visual_flocking/vf_simulation/vf_sims.py",
    "abm/projects/cooperative_signaling/__init__.py",
    "abm/projects/cooperative_signaling/cs_agent/__init__.py",
    "abm/projects/cooperative_signaling/cs_agent/cs_supcalc.py",
    # the experiment definitions of BASELINE configs[2] and configs[3] (run unchanged through abm_b200.compat by the tests)
    "abm/data/metaprotocol/experiments/figExp3BN50PatchyCollOcc.py",
    "abm/data/metaprotocol/experiments/VFExp4c.py",
]


def available() -> bool:
    """True when oracle/_ref holds the copied modules."""
    return os.path.isfile(os.path.join(REF_DST, "abm", "agent", "agent.py"))


def build(verbose: bool = False) -> bool:
    """Copy the files above from the mounted reference (only those that differ).  Returns available()."""
    if not os.path.isdir(os.path.join(REF_SRC, "abm", "agent")):
        return available()
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        if not os.path.isfile(src):
            continue                      # optional package markers
        if os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        if verbose:
            print("[oracle/_ref]", rel)
    return available()


if __name__ == "__main__":
    print("oracle/_ref available:", build(verbose=True))

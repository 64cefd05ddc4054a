#!/usr/bin/env python
"""Builds one libmaven_sm100 variant per hypothesis from gemm_tc_two_group_epilogue.cu.txt (build container, no GPU needed):

    python scripts/experimental/build_variants.py          ->  scripts/experimental/lib_<name>.so   (git-ignored, travel with gpurun)
    gpurun -- 'bash scripts/experimental/run_variants.sh'  ->  failures out of 2*DIAG_REPS runs per variant (scripts/tc_diag.py)

The product source csrc/gemm_tc.cu is swapped only for the duration of each build and restored afterwards."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "multimodal-supernovae_b200", "csrc", "gemm_tc.cu")
EXP = os.path.join(ROOT, "scripts", "experimental", "gemm_tc_two_group_epilogue.cu.txt")
LIB = os.path.join(ROOT, "multimodal-supernovae_b200", "libmaven_sm100.so")
VARIANTS = {
    "split_aux": ["EXP_SPLIT_AUX"],                                   # reproduces the ~40 % failure rate
    "split_aux_delay": ["EXP_SPLIT_AUX", "EXP_DELAY_FIRST_FILL"],
    "split_aux_fence": ["EXP_SPLIT_AUX", "EXP_FENCE_AFTER_WAIT"],
    "split_aux_after_a": ["EXP_SPLIT_AUX", "EXP_AUX_AFTER_A"],
    "split_aux_reread": ["EXP_SPLIT_AUX", "EXP_REREAD"],              # marks rows whose slot content changed after the wait returned
    "split_aux_swap": ["EXP_SPLIT_AUX", "EXP_SWAP_GROUPS"],
    "split_aux_fence_release": ["EXP_SPLIT_AUX", "EXP_FENCE_BEFORE_RELEASE"],   # the fix: reads ordered before the release
    "split_aux_dep_arrive": ["EXP_SPLIT_AUX", "EXP_DEP_ARRIVE"],      # release the slot only after the row loads have returned
    "split_aux_delay_after_release": ["EXP_SPLIT_AUX", "EXP_DELAY_AFTER_RELEASE"],   # early refill, late output: which side matters?           # group 1 reads slot 0 first: does the failure follow the group or the slot?
}
if os.environ.get("EXP_ONLY"):
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in os.environ["EXP_ONLY"].split(",")}


def main():
    ship = open(SRC).read()
    exp = open(EXP).read()
    try:
        for name, defs in VARIANTS.items():
            head = "".join(f"#define {d} 1\n" for d in defs)
            # the round-1 file predates tc_trunc_comp() (ffn_fused.cu links against it): the variants run without the compensation
            open(SRC, "w").write(head + exp + "\nnamespace mvn { float tc_trunc_comp() { return 1.0f; } }\n")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py")], capture_output=True, text=True)
            if "built" not in r.stdout:
                sys.stderr.write(r.stderr[-2000:])
                raise SystemExit(f"variant {name} failed to build")
            shutil.copy(LIB, os.path.join(ROOT, "scripts", "experimental", f"lib_{name}.so"))
            print("built", name)
    finally:
        open(SRC, "w").write(ship)
        subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py")], capture_output=True, text=True)
    print("restored the product library")


if __name__ == "__main__":
    main()

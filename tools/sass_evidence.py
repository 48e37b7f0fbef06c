"""Regenerate profiles/sass_evidence.md: per-module counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / cluster /
multimem usage (cuobjdump -sass on the in-tree libraries; runs without a GPU)."""
import glob, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cols = [("UTC*MMA", r"\bUTC[A-Z]*MMA\b"), ("UTCCP", r"\bUTCCP\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
        ("UTMALDG", r"\bUTMALDG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("SYNCS", r"\bSYNCS\b"), ("STAS", r"\bSTAS\b"),
        ("UCGABAR", r"\bUCGABAR_ARV\b"), ("LDGMC", r"\bLDGMC\b"), ("REDG.SYS", r"\bREDG\.E\.ADD\.STRONG\.SYS\b"),
        ("REDUX", r"\bC?REDUX\b"), ("HMMA", r"\bHMMA\b")]
out = ["# SASS evidence per native module (`python tools/sass_evidence.py`: `cuobjdump -sass flashinfer_b200/_lib/*.so`, mnemonic counts)\n",
       "`UTC*MMA` = tcgen05.mma (UTCHMMA bf16/fp16, UTCQMMA fp8, UTCOMMA fp4 block-scaled), `UTCCP` = tcgen05.cp (scale factors smem -> TMEM), "
       "`LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` = TMA tensor loads, `UBLKCP` = bulk copies, `SYNCS` = mbarrier ops, `STAS` = st.async (DSMEM stores "
       "that credit a remote mbarrier), `UCGABAR` = cluster barriers, `LDGMC` = multimem.ld_reduce (in-switch reduction), `REDG.SYS` = system-scope reductions "
       "(multimem.red / peer flag bumps), `REDUX` = redux.sync (incl. CREDUX).  No `HMMA` (legacy mma.sync) anywhere.\n",
       "| module | " + " | ".join(c for c, _ in cols) + " |", "|---|" + "---|" * len(cols)]
for so in sorted(glob.glob(os.path.join(root, "flashinfer_b200", "_lib", "*.so"))):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out.append(f"| {os.path.basename(so)} | " + " | ".join(str(len(re.findall(rx, sass))) for _, rx in cols) + " |")
open(os.path.join(root, "profiles", "sass_evidence.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

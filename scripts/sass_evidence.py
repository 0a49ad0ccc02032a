#!/usr/bin/env python
"""SASS evidence that the tensor-core path issues tcgen05 / TMA instructions and that k_conn's source stream is a TMA bulk copy:
   python scripts/sass_evidence.py > profiles/r02_sass_tensor_tma.md      (after python -m midaspom_b200.build)"""
import collections, re, subprocess
from pathlib import Path
OBJ = Path(__file__).resolve().parents[1] / "midaspom_b200" / "lib" / "obj"
PAT = re.compile(r"\b(UTCHMMA[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTCATOMSWS[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*)")
print("# SASS evidence (cuobjdump -sass of the sm_100a objects in midaspom_b200/lib/obj)\n")
print("Mnemonics per kernel: `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld (TMEM -> registers), `UTCBAR` = tcgen05.commit, `UTCATOMSWS` = tcgen05.alloc / dealloc,")
print("`UTMALDG` = cp.async.bulk.tensor (TMA tile load), `UBLKCP` = cp.async.bulk (TMA bulk copy), `SYNCS` = mbarrier operations.\n")
for obj in ("mp_conn_gemm.o", "mp_engine.o"):
    sass = subprocess.run(["cuobjdump", "-sass", str(OBJ / obj)], capture_output=True, text=True).stdout
    name, per, sample = None, collections.defaultdict(collections.Counter), {}
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        for op in PAT.findall(ln):
            per[name][op] += 1
            sample.setdefault((name, op.split(".")[0]), ln.strip())
    print(f"## {obj}\n")
    hot = [k for k, cnt in per.items() if any(o.startswith(("UTC", "UTMA", "UBLKCP", "LDTM")) for o in cnt)]
    if len(hot) > 6:
        print(f"{len(hot)} kernel instantiations carry these instructions (every k_conn<precision, geometry, years, culled, targets, threads>); the first is listed.\n")
    for k, cnt in per.items():
        if k not in hot[:5]:
            continue
        print(f"### `{k}`\n")
        print("| mnemonic | static count |\n|---|---|")
        for o, c in sorted(cnt.items()):
            print(f"| {o} | {c} |")
        print("\nfirst occurrence of each:\n\n```")
        for (kk, base), ln in sample.items():
            if kk == k and base.startswith(("UTC", "UTMA", "UBLKCP", "LDTM")):
                print(ln[:150])
        print("```\n")

# the FP32 year contraction of k_conn (mp_conn32.o): packed FP32 FMAs, two years per instruction
sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN2mp6k_connIfLi1ELi20ELb1ELi2ELi128ELb1EEEvNS_8ConnArgsIT_EE", str(OBJ / "mp_conn32.o")], capture_output=True, text=True).stdout
ops = collections.Counter(re.sub(r"\..*", "", m.group(1)) for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", sass, re.M))
first = next((ln.strip() for ln in sass.splitlines() if "FFMA2" in ln), "")
print("## mp_conn32.o\n")
print("### `void mp::k_conn<float, 1, 20, true, 2, 128, true>` (cfg3: coordinates, 19 years, culled, 128 threads x 2 targets, FP32 contraction)\n")
print("| mnemonic | static count |\n|---|---|")
for o in ("FFMA2", "FFMA", "DFMA", "DADD", "F2F", "MUFU", "UBLKCP", "LDS", "STS"):
    print(f"| {o} | {ops.get(o, 0)} |")
print("\n`FFMA2` = fma.rn.f32x2 (the weight as a broadcast `.F32` operand, two years of the 0/1 table as the `.F32x2` operand); no DFMA: the FP64 work is the join of the partial sums (`F2F` + `DADD`).\n\n```")
print(first[:150])
print("```")

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

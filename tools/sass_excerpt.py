"""SASS evidence that the shipped library runs tcgen05 / TMA / TMEM code (B200_PROFILING.md: UTCHMMA = tcgen05.mma,
UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKPF = bulk L2 prefetch).

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.md     (needs cuobjdump; run where the .so was built)"""
import collections
import re
import subprocess
import sys

LIB = "stylesinger_b200/libstylesinger_b200.so"
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKPF", "LDTM", "UTCBAR", "UTCBAR.2CTA", "UTCATOMSWS", "SYNCS.ARRIVE", "SYNCS.PHASECHK"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn) + r"\b", line):
                per[cur][mn] += 1
    demangled = subprocess.run(["cu++filt"] + list(per.keys()), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(per.keys(), demangled)) if len(demangled) == len(per) else {k: k for k in per}
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("# Round 2 - SASS excerpt of libstylesinger_b200.so (cuobjdump -sass, sm_100a)\n")
    print("Totals over the library: " + ", ".join(f"{mn} {tot[mn]}" for mn in MNEMONICS) + ".\n")
    print("Per kernel (only kernels that contain tensor-core / TMA instructions):\n")
    print("| kernel | " + " | ".join(MNEMONICS) + " |\n|---|" + "---:|" * len(MNEMONICS))
    for k, c in per.items():
        if c["UTCHMMA"] or c["UTMALDG"]:
            short = names[k].split(">(")[0] + (">" if ">(" in names[k] else "")
            short = short.replace("(int)", "").replace("void ", "").replace("ssb::<unnamed>::", "").split("(CUtensorMap")[0]
            print(f"| `{short}` | " + " | ".join(str(c[mn]) for mn in MNEMONICS) + " |")
    # a short literal excerpt of the tap-reuse kernel's MMA issue loop
    print("\nFirst tcgen05.mma group of `conv_gemm_tc2r_kernel<128, GATE>` (3 products per K step: hi*hi, hi*lo, lo*hi):\n\n```")
    on, shown = False, 0
    for line in sass.splitlines():
        if "Function :" in line:
            on = "conv_gemm_tc2r_kernelILi128ELi1" in line
        if on and ("UTCHMMA" in line or "UTCBAR" in line) and shown < 8:
            print(re.sub(r"/\*[0-9a-f]{4}\*/", "", line).strip()[:150])
            shown += 1
    print("```")


if __name__ == "__main__":
    sys.exit(main())

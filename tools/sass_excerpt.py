"""per-kernel SASS evidence for profiles/ (no GPU needed): counts of the tcgen05 / TMA / cluster mnemonics in every kernel of
libmpn_b200.so + a short excerpt around the first UTCHMMA of the dominant kernels.
   python tools/sass_excerpt.py > profiles/r02_sass_excerpt.md"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "multipathnet_b200/libmpn_b200.so"
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UCGABAR", "FFMA", "HMMA", "LDG.E.128", "STG.E.128", "LDS", "FMNMX", "ATOMS"]
kernels = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1); kernels[name] = []
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        kernels[name].append(line)
dem = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence per kernel of `%s` (`cuobjdump -sass`, sm_100a)\n" % LIB)
print("Counts of static instructions. `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` = TMA tensor "
      "load / store, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `UCGABAR` = cluster barrier. A tensor-core kernel with 0 `FFMA` in its "
      "main loop does its contraction on the tensor pipe only.\n")
print("| kernel | instr | " + " | ".join(WANT) + " |")
print("|---|---|" + "---|" * len(WANT))
for (k, lines), d in zip(kernels.items(), dem):
    short = re.sub(r"\(anonymous namespace\)::", "", d)
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    text = "\n".join(lines)
    cnt = []
    for w in WANT:
        if w == "UTCHMMA":
            cnt.append(len(re.findall(r"\bUTCHMMA\b(?!\.2CTA)", text)))
        else:
            cnt.append(len(re.findall(r"\b" + re.escape(w) + r"\b", text)))
    print(f"| `{short[:70]}` | {len(lines)} | " + " | ".join(str(c) for c in cnt) + " |")
print()
for pat, title, mn in (("conv_gemm_tc_kernel<240, 2, true>", "fc6 / fc7 (fp16 x fp16 two-product kernel, CTA pairs)", "UTCHMMA"),
                       ("conv3x3_tc_kernel<256, 2>", "3x3 trunk convolution (A-reuse kernel, CTA pairs)", "UTCHMMA"),
                       ("conv3x3_tc_kernel<256, 2>", "3x3 trunk convolution: first TMEM read of the epilogue", "LDTM"),
                       ("roi_pool_cluster_kernel", "fused Foveal + ROI pooling (4-CTA clusters)", "UCGABAR")):
    for (k, lines), d in zip(kernels.items(), dem):
        if pat in d:
            idx = next((i for i, l in enumerate(lines) if mn in l), 0)
            print(f"## {title}: `{pat}`, instructions {max(0, idx - 6)}..{idx + 10}\n\n```")
            for l in lines[max(0, idx - 6): idx + 10]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip())
            print("```\n")
            break

"""Per-kernel SASS mnemonic counts of the built library (no GPU needed): the evidence that the kernels use the
Blackwell paths DESIGN.md names.  usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "handposeestimation-with-3d-cnns_b200", "lib", "obj")
WANT = ["UBLKCP.S.G", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "FFMA2", "FMUL2", "FADD2", "CREDUX.MIN.F32",
        "CREDUX.MAX.F32", "STG.E.128", "LDS.128", "ATOMS.MIN.S32", "ATOMS.MAX.S32", "ATOMS.OR", "DMUL", "MUFU.RCP64H",
        "REDUX.SUM.S32", "FMNMX3", "MUFU.RCP", "FCHK", "HMMA", "UTCHMMA", "STL", "LDL"]


def demangle(name):
    try:
        out = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        out = name
    out = re.sub(r"dtsdf::\(anonymous namespace\)::|dtsdf::<unnamed>::|void ", "", out)
    return out[:out.rindex(">(") + 1] if ">(" in out else out.split("(")[0]


print("# SASS mnemonics per kernel of lib/libdtsdf_b200.so: cuobjdump -sass of the objects __graft_entry__.build() links\n"
      "# (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false), counted by tools/sass_evidence.py.\n"
      "# UBLKCP.S.G = cp.async.bulk global->shared (TMA bulk copy); SYNCS.*TRANS64 = mbarrier expect-tx / try_wait;\n"
      "# FFMA2/FMUL2/FADD2 = Blackwell packed FP32 pipe; CREDUX = redux.sync.{min,max}.f32; ATOMS.OR / DMUL = the sampled bounding\n"
      "# box; no HMMA / UTC*MMA anywhere (there is no contraction on this path); STL/LDL = spills.\n"
      "# FMNMX3 = three-input min / max (sm_100); MUFU.RCP against FCHK: every IEEE division carries one of each, a reciprocal\n"
      "# sequence without range check (the augmented kernel's safe-frame division, the fixed-divisor forms) only the former.\n"
      "# dtsdf_fused_kernel<R, LAYOUT, DEFF, SAMPLED>, dtsdf_aug_kernel<R, LAYOUT>.")
for obj in ("dtsdf_kernels.cu.o", "dtsdf_codec.cu.o"):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    name, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            counts[name][m.group(1)] += 1
    for n, c in counts.items():
        agg = collections.Counter()
        for op, k in c.items():
            for w in WANT:
                if op.startswith(w):
                    agg[w] += k
        print("%-62s %6d instr  %s" % (demangle(n)[:62], sum(c.values()), "  ".join("%s=%d" % (w, agg[w]) for w in WANT if agg[w])))

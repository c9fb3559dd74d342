"""Committed SASS evidence: `cuobjdump -sass` listing of every hot kernel of the shipped extension,
one file per kernel under profiles/sass/, plus a mnemonic summary (UTC*MMA = tcgen05.mma, `.2CTA`
= cta_group::2, UTMALDG = TMA, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, LDGMC = multimem
ld_reduce, SYNCS = mbarrier, FFMA2 = packed fp32x2; /opt/skills/guides/B200_PROFILING.md).

  python tools/dump_sass.py            # needs only the built .so and cuobjdump (no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tensorflowonspark_b200", "_ext", "_tfos_b200_C.so")
OUT = os.path.join(ROOT, "profiles", "sass")

# (file stem, regex on the demangled name)
KERNELS = [
    ("igemm_fwd_256_fprop_stats_2cta", r"igemm_fwd_kernel<256, false, 1, 2>"),
    ("igemm_fwd_256_fprop_stats_1cta", r"igemm_fwd_kernel<256, false, 1, 1>"),
    ("igemm_fwd_256_dgrad_accum_bnred_1cta", r"igemm_fwd_kernel<256, true, 10, 1>"),
    ("igemm_fwd_256_dgrad_bnred_2cta", r"igemm_fwd_kernel<256, true, 8, 2>"),
    ("igemm_fwd_64_fprop_stats", r"igemm_fwd_kernel<64, false, 1, 1>"),
    ("igemm_wgrad_128", r"igemm_wgrad_kernel<128>"),
    ("igemm_wgrad_wide", r"igemm_wgrad_wide_kernel"),
    ("igemm_wgrad_halo", r"igemm_wgrad_halo_kernel"),
    ("allreduce_opt_momentum_nvls", r"allreduce_opt_kernel<1, true, 0>"),
    ("allreduce_opt_momentum_p2p", r"allreduce_opt_kernel<1, false, 0>"),
    ("allreduce_phase1_reduce_scatter", r"allreduce_opt_kernel<0, false, 1>"),
    ("allreduce_phase2_momentum_update_allgather", r"allreduce_opt_kernel<1, false, 2>"),
    ("ps_apply_momentum", r"ps_apply_kernel<1>"),
    ("ps_push_slot", r"ps_push_slot_kernel"),
    ("ps_pull_model", r"ps_pull_model_kernel"),
    ("stem_bn_relu_pool_fwd", r"stem_bn_relu_pool_fwd_kernel"),
    ("stem_pool_bn_bwd_reduce", r"stem_pool_bn_bwd_reduce_kernel"),
    ("stem_pool_bn_bwd_apply", r"stem_pool_bn_bwd_apply_kernel"),
    ("igemm_fwd_256_generic_bias_relu_accumulate", r"igemm_fwd_kernel<256, false, 4, 1>"),
    ("bn_fwd_apply_finalize", r"bn_fwd_apply_kernel<true>"),
    ("bn_bwd_apply", r"bn_bwd_apply_kernel"),
    ("bn_bwd_reduce", r"bn_bwd_reduce_kernel"),
]
KEY = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCCP", "LDGMC", "SYNCS", "FFMA2", "HMMA", "RED", "ATOM",
       "UCGABAR", "ST.E", "STG", "LDG")


def main():
  text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
  funcs = collections.OrderedDict()
  cur = None
  for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
      cur = m.group(1)
      funcs[cur] = []
      continue
    if cur is not None:
      funcs[cur].append(line)
  names = list(funcs)
  dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
  os.makedirs(OUT, exist_ok=True)
  summary = ["# mnemonic counts per kernel (tools/dump_sass.py; full listings next to this file)", ""]
  for stem, pat in KERNELS:
    hit = [n for n, d in zip(names, dem) if re.search(re.escape(pat), d)]
    if not hit:
      summary.append("{:42s} NOT FOUND ({})".format(stem, pat))
      continue
    body = funcs[hit[0]]
    # drop the encoding column: the mnemonics are the evidence, the hex doubles the size
    clean = [re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", l) for l in body]
    clean = [l for l in clean if l.strip() and not re.match(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", l)]
    with open(os.path.join(OUT, stem + ".sass"), "w") as f:
      f.write("// {}\n// {}\n".format(dem[names.index(hit[0])], hit[0]))
      f.write("\n".join(clean) + "\n")
    ops = collections.Counter()
    for l in clean:
      m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
      if m:
        ops[m.group(1)] += 1
    total = sum(ops.values())
    keyed = collections.Counter()
    for op, c in ops.items():
      for k in KEY:
        if op.startswith(k):
          keyed[op] += c
    summary.append("{} ({} instructions)".format(stem, total))
    for op, c in sorted(keyed.items(), key=lambda kv: (-kv[1], kv[0]))[:18]:
      summary.append("    {:34s} {}".format(op, c))
    summary.append("")
  with open(os.path.join(OUT, "SUMMARY.txt"), "w") as f:
    f.write("\n".join(summary) + "\n")
  print("\n".join(summary[:60]))


if __name__ == "__main__":
  sys.exit(main())

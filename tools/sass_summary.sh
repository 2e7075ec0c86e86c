#!/bin/bash
# SASS evidence for profiles/: per-kernel instruction histogram of the mnemonics that prove the hardware path (tcgen05 MMA / TMEM /
# TMA / bulk copies / packed fp32 / warp-level MMA), plus an excerpt of one MMA issue sequence.   usage: tools/sass_summary.sh OUTDIR
cd "$(dirname "$0")/.."
OUT=${1:-profiles/r02}
mkdir -p $OUT
{
echo "# cuobjdump -sass of fastenhancer_b200/_build/fe_inst_16b.o (FastEnhancer_B, 16 kHz), one line per kernel variant"
echo "# Plan<Cfg<n_fft,hop,C1,E,C2,F2,K,heads>, streams per CTA, family> -- family 0 fp32 FMA pipe, 1 TF32, 2 fp16, 3 bf16 conv section, 4 split-fp16 (fp32x3)"
echo "# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = cp.async.bulk.tensor (2-D TMA tiles),"
echo "# UBLKCP = cp.async.bulk (weight ring), HMMA.1688.F32.TF32 = mma.sync m16n8k8 (attention), FFMA2 / FADD2 / FMUL2 = packed fp32, MUFU = SFU"
cuobjdump -sass fastenhancer_b200/_build/fe_inst_16b.o | awk '
/Function :/ { name=$3; if (match(name, /EEELi[0-9]+ELi[0-9]+EEEEE/)) { t=substr(name, RSTART+5, RLENGTH-10); split(t, q, "ELi"); name=sprintf("Plan<C16B, S=%s, family=%s>", q[1], q[2]); } else name="(other) " substr(name,1,40); }
/^ +\/\*[0-9a-f]+\*\/ / { total[name]++; m=$2; if (m ~ /^@/) m=$3;
   if (m ~ /^UTCHMMA/) c[name,"UTCHMMA"]++; else if (m ~ /^LDTM/) c[name,"LDTM"]++; else if (m ~ /^STTM/) c[name,"STTM"]++;
   else if (m ~ /^UTCBAR/) c[name,"UTCBAR"]++; else if (m ~ /^UTMALDG/) c[name,"UTMALDG"]++; else if (m ~ /^UTMASTG/) c[name,"UTMASTG"]++;
   else if (m ~ /^UBLKCP/) c[name,"UBLKCP"]++; else if (m ~ /^HMMA/) c[name,"HMMA"]++; else if (m ~ /^FFMA2/) c[name,"FFMA2"]++;
   else if (m ~ /^FADD2|^FMUL2/) c[name,"FADD2/FMUL2"]++; else if (m ~ /^MUFU/) c[name,"MUFU"]++; else if (m ~ /^FFMA/) c[name,"FFMA"]++;
   else if (m ~ /^SYNCS/) c[name,"SYNCS"]++; else if (m ~ /^BAR/) c[name,"BAR"]++; }
END { split("UTCHMMA LDTM STTM UTCBAR UTMALDG UTMASTG UBLKCP HMMA FFMA2 FADD2/FMUL2 FFMA MUFU SYNCS BAR", k, " ");
  for (n in total) { line = sprintf("%-32s total %6d", n, total[n]); for (i = 1; i <= 14; i++) line = line sprintf("  %s %d", k[i], c[n,k[i]] + 0); print line } }' | sort
} > $OUT/sass_histogram_16k_b.txt
# an excerpt: the first tcgen05.mma issue sequence of the fp32-accurate B kernel (2 streams per CTA) and the TMA tile instructions
cuobjdump -sass fastenhancer_b200/_build/fe_inst_16b.o | awk '/Function :.*Li2ELi4EEEEE/ {on=1} on && /Function :/ && !/Li2ELi4EEEEE/ {on=0} on' > /tmp/sass_b_x3.txt
{
echo "# fe_fused_kernel<Plan<C16B, 2, 4>> (fp32-accurate family): first tcgen05.mma issue sequence (elect + predicated UTCHMMA with incremental descriptors)"
grep -n "UTCHMMA" /tmp/sass_b_x3.txt | head -1 | cut -d: -f1 | xargs -I{} awk -v s={} 'NR >= s - 12 && NR <= s + 40' /tmp/sass_b_x3.txt
echo
echo "# ... and its 2-D TMA tile instructions (hop tile in / out) with the surrounding mbarrier / fence code"
grep -n "UTMALDG\|UTMASTG" /tmp/sass_b_x3.txt | head -2 | cut -d: -f1 | while read l; do awk -v s=$l 'NR >= s - 6 && NR <= s + 4' /tmp/sass_b_x3.txt; echo "  ..."; done
} > $OUT/sass_excerpt_16k_b_fp32x3.txt
wc -l $OUT/sass_histogram_16k_b.txt $OUT/sass_excerpt_16k_b_fp32x3.txt

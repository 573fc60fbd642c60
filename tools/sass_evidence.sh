#!/bin/bash
# SASS evidence that the tensor-core kernels are Blackwell-native: per kernel of libfdm_b200.so, the count of tcgen05 MMA
# (UTCHMMA), TMEM load (LDTM), TMA load / store (UTMALDG / UTMASTG, UBLKCP for 1-D bulk copies), commit barrier (UTCBAR) and,
# for contrast, legacy mma.sync (HMMA) instructions. Runs without a GPU:  bash tools/sass_evidence.sh > profiles/r02_sass_evidence.md
SO=face-diffusion-model_b200/fdm_b200/libfdm_b200.so
echo "# SASS instruction counts per kernel of libfdm_b200.so (cuobjdump -sass, sm_100a)"
echo
echo "| kernel | UTCHMMA | LDTM | UTMALDG | UTMASTG | UBLKCP | UTCBAR | HMMA (mma.sync) |"
echo "|---|---|---|---|---|---|---|---|"
cuobjdump -sass $SO | awk '
/Function :/ {name=$3}
/UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /UTMASTG/ {d[name]++} /UBLKCP/ {e[name]++} /UTCBAR/ {f[name]++}
/ HMMA\./ {g[name]++}
END { for (n in a) k[n]=1; for (n in g) k[n]=1; for (n in c) k[n]=1;
      for (n in k) printf "%s %d %d %d %d %d %d %d\n", n, a[n], b[n], c[n], d[n], e[n], f[n], g[n] }' | while read n a b c d e f g; do
  echo "| \`$(echo $n | c++filt | sed 's/(anonymous namespace):://; s/void //; s/(.*//' | cut -c1-90)\` | $a | $b | $c | $d | $e | $f | $g |"
done | sort

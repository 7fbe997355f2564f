#!/bin/sh
# Evidence that the shipped library runs Blackwell-native instructions: per-kernel counts of the tcgen05 / TMA SASS
# mnemonics (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA bulk-tensor load / store, LDTM = tcgen05.ld, UTCBAR =
# tcgen05.commit) in tgan/libtgan.so, plus the first occurrences in igemm_kernel / wgrad_kernel.
#   sh tools/sass_excerpt.sh > profiles/r2_sass_excerpt.txt
SO=tensorflow-implementation-of-triple-gan_b200/tgan/libtgan.so
cuobjdump -sass "$SO" > /tmp/tgan_sass.txt
echo "# cuobjdump -sass $SO  ($(cuobjdump -lelf "$SO" | head -1))"
awk '/Function :/ {fn=$3} /UTCHMMA|UTMALDG|UTMASTG|LDTM|UTCBAR|UTMAPF|SYNCS/ {split($0,a," "); for(i in a) if (a[i] ~ /^(UTCHMMA|UTMALDG|UTMASTG|LDTM|UTCBAR|UTMAPF|SYNCS)/) {sub(/\..*/,"",a[i]); c[fn" "a[i]]++}} END {for (k in c) print c[k], k}' /tmp/tgan_sass.txt | sort -k2,2 -k3,3 | awk '{printf "%6d  %-60s %s\n", $1, $2, $3}'
for k in igemm_kernel wgrad_kernel; do
  echo "# ---- first tcgen05 / TMA instructions of $k"
  awk -v k="$k" '/Function :/ {on = index($0, k) > 0} on && /UTCHMMA|UTMALDG|UTMASTG|LDTM|UTCBAR/' /tmp/tgan_sass.txt | head -12
done

# A/B of library builds: bash scripts/ab.sh "<variant> ..." "<command>"   (variant v => vectorindex_b200/libvindex_v.so)
for V in $1; do
echo "== $V"; VIX_LIB_PATH=$PWD/vectorindex_b200/libvindex_$V.so $2 2>&1 | tail -1 | cut -c1-330
done

#!/bin/bash
# time the spectral step with every library variant under build_mb/lib_*.so (kernel experiments; see profiles/r2_spectral64.md)
cp pyqg_generative_b200/libqgb200.so /tmp/lib_orig.so
for f in build_mb/lib_*.so; do
  cp $f pyqg_generative_b200/libqgb200.so
  echo "== $f"; python scripts/spectral_time.py ${1:-64} ${2:-1024} 200
done
cp /tmp/lib_orig.so pyqg_generative_b200/libqgb200.so

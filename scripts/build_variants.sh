# builds libmld_cuda variants with different compile-time tunables into gpurun_out/variants/
set -e
cd "$(dirname "$0")/../mono_lidar_depth_b200/csrc"
mkdir -p ../../build/variants
build() { # name, extra flags
  rm -f *.o
  make -j8 EXTRA="$2" OUT=../../build/variants/libmld_$1.so > /dev/null 2>&1 || { echo "build $1 failed"; exit 1; }
  echo built $1
}
build tbt64 "-DMLD_T_TBT=64"
build tbt256 "-DMLD_T_TBT=256"
build tcap8 "-DMLD_T_TCAP=8"
build tcap8_tbt64 "-DMLD_T_TCAP=8 -DMLD_T_TBT=64"
build ppt8 "-DMLD_K1_PPT=8"
build ppt2 "-DMLD_K1_PPT=2"
build k1t128 "-DMLD_K1_THREADS=128"
rm -f *.o
make -j8 > /dev/null 2>&1
echo done

# builds libmld_cuda variants with different compile-time tunables into build/variants/ (travels to the GPU box)
set -e
cd "$(dirname "$0")/../mono_lidar_depth_b200/csrc"
mkdir -p ../../build/variants
build() { # name, extra flags
  rm -f *.o
  make -j8 EXTRA="$2" OUT=../../build/variants/libmld_$1.so > /dev/null 2>&1 || { echo "build $1 failed"; exit 1; }
  echo built $1
}
for spec in "$@"; do build "${spec%%:*}" "${spec#*:}"; done
rm -f *.o
make -j8 > /dev/null 2>&1
echo done

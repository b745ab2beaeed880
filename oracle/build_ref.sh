#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY — builds the UNMODIFIED reference (FASP 2.8.7) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored).
#
#   oracle/_ref/libfasp_seq.so  sequential build  = the parity oracle (SURVEY.md finding 2)
#   oracle/_ref/libfasp_omp.so  OpenMP build      = CPU timing baseline only (runs multicolour GS
#                                                   regardless of the requested smoother)
#
# The reference's own build system (cmake) is NOT run; this is a direct gcc recipe over
# base/src/*.c + base/extra/{hb_io,interface}/*.c (no third-party libs: all WITH_* default 0).
# The OpenMP variant needs a one-token fix (BlaSpmvBSR.c:56 `A->nnz` -> `A->NNZ`) which is
# applied to a private copy under a temp dir; nothing is copied into this repository.
set -euo pipefail
REF=${FASP_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/base/src" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
INC="-I$REF/base/include -I$REF/base/extra/include"
SRCS="$(ls $REF/base/src/*.c $REF/base/extra/hb_io/*.c $REF/base/extra/interface/*.c)"
NPROC=$(nproc)

build_one () {  # $1 = name, $2 = extra flags, $3 = "patch" or ""
  local name=$1 flags=$2 patch=$3
  local tmp; tmp=$(mktemp -d /tmp/fasp_ref_${name}.XXXXXX)
  local list="$SRCS"
  if [ -n "$patch" ]; then
    sed 's/A->nnz > OPENMP_HOLDS/A->NNZ > OPENMP_HOLDS/' "$REF/base/src/BlaSpmvBSR.c" > "$tmp/BlaSpmvBSR.c"
    list="$(echo "$SRCS" | grep -v '/BlaSpmvBSR.c$') $tmp/BlaSpmvBSR.c"
  fi
  ( cd "$tmp" && echo $list | tr ' ' '\n' | xargs -P"$NPROC" -n4 \
      gcc -O3 -std=gnu99 -fPIC -w $flags -I"$REF/base/src" $INC -c )
  gcc -shared $flags -o "$OUT/libfasp_${name}.so" "$tmp"/*.o -lm
  rm -rf "$tmp"
  echo "built $OUT/libfasp_${name}.so"
}

if [ ! -f "$OUT/libfasp_seq.so" ] || [ "${1:-}" = "--force" ]; then build_one seq "" ""; fi
if [ ! -f "$OUT/libfasp_omp.so" ] || [ "${1:-}" = "--force" ]; then build_one omp "-fopenmp" patch; fi

# reference-arm bench driver (our code, reference headers + libfasp_omp.so)
if [ ! -f "$OUT/fasp_ref_bench" ] || [ "$HERE/ref_bench.c" -nt "$OUT/fasp_ref_bench" ] || [ "${1:-}" = "--force" ]; then
  gcc -O2 -std=gnu99 -fopenmp -w $INC "$HERE/ref_bench.c" -o "$OUT/fasp_ref_bench" \
      -L"$OUT" -lfasp_omp -lm -Wl,-rpath,'$ORIGIN'
  echo "built $OUT/fasp_ref_bench"
fi

#!/bin/bash
# tools/try_variants.sh -- build kernel variants here (no GPU), measure them all in ONE gpurun call.
#
#   tools/try_variants.sh build  "u8:-DFSK_UNROLL=8"  "ring768::FSK_B200_RING=768"  "noearly:-DFSK_NO_EARLY_REQ"
#       each spec is  name:NVCC_FLAGS[:ENV=VAL ...]  ; empty flags = the in-tree library.
#       Builds variants/lib_<name>.so (git-ignored, shipped to the GPU box) and writes variants/run.sh.
#       Check a variant's logic first on the emulator:
#         make -C tests/emu OUT=$PWD/variants/emu_<name>.so BUILD=$PWD/variants/emubuild_<name> EXTRA="<flags>"
#         FSK_B200_EMU=1 FSK_EMU_LIB=$PWD/variants/emu_<name>.so python -m pytest tests/test_gpu_parity.py -m gpu -k "not roundtrip"
#   gpurun --timeout 900 -- 'bash variants/run.sh'
#       prints one line per variant: Msamples/s, kernel ms, roofline fraction (device-resident bench, 3 steps).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
[ "$1" = build ] || { sed -n 2,14p "$0"; exit 1; }
shift
mkdir -p "$ROOT/variants"
RUN="$ROOT/variants/run.sh"
cat > "$RUN" <<'EOS'
#!/bin/bash
one() { label=$1; lib=$2; shift 2; env FSK_B200_LIB=$lib "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$label:', round(d['value']), 'Ms/s', round(d['roofline']['kernel_ms'],2), 'ms frac', round(d['roofline']['frac'],3))"; }
EOS
for spec in "$@"; do
    name=${spec%%:*}; rest=${spec#*:}; flags=${rest%%:*}; envs=""
    [ "$rest" != "$flags" ] && envs=${rest#*:}
    if [ -n "$flags" ]; then
	make -s -C "$ROOT/minimodem_b200/csrc" OUT="$ROOT/variants/lib_$name.so" BUILD="$ROOT/variants/build_$name" EXTRA="$flags"
	lib="/root/repo/variants/lib_$name.so"
    else
	lib="/root/repo/minimodem_b200/libfsk_b200.so"
    fi
    echo "one \"$name\" $lib $envs" >> "$RUN"
done
echo "wrote $RUN:"; tail -n +4 "$RUN"

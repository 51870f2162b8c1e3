#!/bin/bash
# Count the SASS mnemonics that prove the Blackwell-native paths, per object file of libavb200.so.
#   bash tools/sass_opcodes.sh > profiles/r2_sass_opcodes.txt
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass <object> | grep -c <mnemonic>   ($(nvcc --version | tail -1))"
printf "%-22s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" object UTCHMMA UTCBAR LDTM STTM UTMALDG UBLKCP SYNCS HMMA HFMA2
for o in animal_vision_b200/lib/obj/*.o; do
    s=$(cuobjdump -sass "$o")
    c() { echo "$s" | grep -c "$1"; }
    printf "%-22s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" "$(basename $o)" $(c UTCHMMA) $(c UTCBAR) $(c "LDTM") $(c "STTM") $(c UTMALDG) $(c UBLKCP) $(c "SYNCS") $(c " HMMA") $(c "HFMA2")
done

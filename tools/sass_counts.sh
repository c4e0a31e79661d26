#!/bin/bash
# SASS evidence of the sm_100a instructions in the built library (no GPU needed): tcgen05 MMA (UTCHMMA), TMA loads / stores
# (UTMALDG / UTMASTG), TMEM loads (LDTM), bulk copies (UBLKCP), per object file.
lib=${1:-drone_yolo_b200/lib/libdroneyolo.so}
echo "# $(date -u +%F) $(nvcc --version | tail -2 | head -1)"
for o in drone_yolo_b200/build/*.o; do
  s=$(cuobjdump -sass $o 2>/dev/null)
  printf "%-22s UTCHMMA %5d  UTMALDG %5d  UTMASTG %5d  LDTM %4d  UBLKCP %4d  UTCBAR/commit %4d  kernels %3d\n" $(basename $o) \
    $(echo "$s" | grep -c "UTCHMMA") $(echo "$s" | grep -c "UTMALDG") $(echo "$s" | grep -c "UTMASTG") $(echo "$s" | grep -c "LDTM") \
    $(echo "$s" | grep -c "UBLKCP") $(echo "$s" | grep -c "UTCBAR") $(echo "$s" | grep -c "Function :")
done
echo "arch: $(cuobjdump -lelf $lib 2>/dev/null | grep -o 'sm_[0-9a-z]*' | sort -u | tr '\n' ' ')"

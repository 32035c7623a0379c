#!/bin/bash
# usage: scripts/sass_mix.sh <object.o> <mangled-function-substring>
# instruction mix between the first and the last MUFU.EX2 of a kernel (the epilogue's arithmetic) + per-MUFU count
cuobjdump -sass -fun "$2" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/_f.txt
n=$(wc -l < /tmp/_f.txt); m=$(grep -c 'MUFU.EX2' /tmp/_f.txt)
first=$(grep -n "MUFU.EX2" /tmp/_f.txt | head -1 | cut -d: -f1); last=$(grep -n "MUFU.EX2" /tmp/_f.txt | tail -1 | cut -d: -f1)
echo "$2: $n instructions, $m MUFU.EX2, span $((last-first+1)) -> $(( (last-first+1) / (m>0?m:1) )) per MUFU (static)"
sed -n "${first},${last}p" /tmp/_f.txt | awk '{print $2}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn | head -${3:-22}

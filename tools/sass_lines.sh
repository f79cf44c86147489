#!/bin/bash
# usage: tools/sass_lines.sh <cubin> <kernel-substring> [top]  -- instruction count per source line of one kernel (nvdisasm -g)
nvdisasm -g -c "$1" 2>/dev/null | awk -v want="$2" '
/^[ \t]*\.section[ \t]+\.text\./ {on = index($0, want) > 0}
/\/\/## File/ {match($0,/line [0-9]+/); ln=substr($0,RSTART+5,RLENGTH-5); f=$0; sub(/.*File "/,"",f); sub(/".*/,"",f); n=split(f,a,"/"); cur=a[n]":"ln}
on && /^[ \t]+\/\*[0-9a-f]+\*\// {cnt[cur]++; tot++}
END {for (k in cnt) print cnt[k], k; print tot, "TOTAL"}' | sort -rn | head -${3:-40}

#!/bin/bash
# usage: ncu_lines.sh <report.ncu-rep> <object.o> <kernel substring> [top]: executed warp instructions / stall samples per source line
rep=$1; obj=$(realpath $2); kern=$3; top=${4:-50}
tmp=$(mktemp -d)
ncu -i $rep --page source --csv --print-source sass 2>/dev/null > $tmp/src.csv
( cd $tmp && cuobjdump -xelf all $obj >/dev/null 2>&1 && nvdisasm -g *.cubin > cells.sass 2>/dev/null )
python $(dirname $0)/sass_lines.py $tmp/src.csv $tmp/cells.sass $kern $top
rm -rf $tmp

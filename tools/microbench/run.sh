#!/bin/bash
# Runs the variants of dram_gran.bin in separate processes (plain, then under ncu for DRAM bytes per launch).
cd "$(dirname "$0")"
OUT=../../gpurun_out
mkdir -p $OUT
M=dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_sectors.sum,gpu__time_duration.sum
SEL=${1:-param}
timeout 60 ./dram_gran.bin 0 262144 $SEL > $OUT/gran_plain_$SEL.txt 2>&1
cat $OUT/gran_plain_$SEL.txt
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $OUT/gran_ncu_$SEL.csv ./dram_gran.bin 0 262144 $SEL > $OUT/gran_ncu_$SEL.log 2>&1
tail -2 $OUT/gran_ncu_$SEL.log

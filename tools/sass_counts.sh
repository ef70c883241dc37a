#!/bin/bash
# SASS evidence per kernel of libsmoke_b200.so (run here, no GPU needed): instruction counts that show what the kernels are made of --
# UTMALDG (TMA tensor loads), SYNCS (mbarrier transactions), FFMA2/FADD2/FMUL2 (packed f32x2), F2F + DMUL (the double-precision
# multiply by 1.9, cu:384), LDS/STS, SHFL, BAR.   usage: tools/sass_counts.sh > profiles/r2_sass_counts.txt
cd "$(dirname "$0")/.." || exit 1
SO=smoke-simulation_b200/libsmoke_b200.so
echo "# $(date -u +%F) $(nvcc --version | tail -1)  $SO"
printf "%-62s %6s %7s %6s %6s %6s %5s %5s %5s %5s %5s %5s\n" kernel instr UTMALDG SYNCS FFMA2 FADD2 FMUL2 F2F DMUL LDS STS SHFL
cuobjdump -sass $SO | c++filt | awk '
/Function : /{ if (name != "") flush(); name=substr($0, index($0, "Function : ") + 11); sub(/\(.*/, "", name); sub(/^void /, "", name); n=0; delete c }
/^[ \t]+\/\*[0-9a-f]+\*\//{ n++; op=$2; sub(/\..*/,"",op); if (op ~ /^@/) { op=$3; sub(/\..*/,"",op) } c[op]++ }
function flush() { printf "%-62s %6d %7d %6d %6d %6d %6d %5d %5d %5d %5d %5d\n", substr(name,1,62), n, c["UTMALDG"], c["SYNCS"], c["FFMA2"], c["FADD2"], c["FMUL2"], c["F2F"], c["DMUL"], c["LDS"], c["STS"], c["SHFL"] }
END{ flush() }' | grep -E "k_pressure_reg<4, 16|k_pressure_tma<4|k_advect|k_pressure_half|k_jacobi|k_force|k_fill|k_codes|k_hash|k_absmax|k_margin|k_mask|k_density|k_max_div|k_epoch|k_copy16|^kernel"

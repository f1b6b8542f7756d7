#!/bin/bash
# SASS evidence for profiles/: per kernel, the counts of the instructions that matter on this path
# (DMMA = FP64 tensor pipe; LDGSTS = cp.async; UTMALDG = TMA bulk tensor load; SYNCS = mbarrier; LDS/STS;
# BAR) and the first DMMA / copy lines verbatim.  Usage: tools/sass_excerpt.sh > profiles/r02_sass_excerpts.txt
LIB=multifidelity_datafusion_gps_b200/libmfgp_b200.so
echo "# cuobjdump -sass $LIB (sm_100a), $(date -u +%F)"
echo "# FP64 MMA on sm_100a is DMMA.8x8x4 (tcgen05/UMMA has no f64 kind); operands arrive by cp.async (LDGSTS)"
echo "# or, in trmm_sumsq_tma_kernel, by TMA (UTMALDG.2D) signalling mbarriers (SYNCS.*)."
cuobjdump -sass $LIB 2>/dev/null | awk '
/Function : /{ if (name != "") report(); name=$3; n=0; delete cnt; first_dmma=""; first_cp=""; first_tma=""; first_sy="" }
/\/\*[0-9a-f]+\*\//{
  line=$0
  if (line ~ /DMMA/)   { cnt["DMMA"]++;   if (first_dmma=="") first_dmma=line }
  if (line ~ /LDGSTS/) { cnt["LDGSTS"]++; if (first_cp=="") first_cp=line }
  if (line ~ /UTMALDG/){ cnt["UTMALDG"]++; if (first_tma=="") first_tma=line }
  if (line ~ /SYNCS/)  { cnt["SYNCS"]++;  if (first_sy=="") first_sy=line }
  if (line ~ / LDS/)   cnt["LDS"]++
  if (line ~ / STS/)   cnt["STS"]++
  if (line ~ /BAR\.SYNC/) cnt["BAR.SYNC"]++
  if (line ~ /DFMA/)   cnt["DFMA"]++
  if (line ~ /DMUL/)   cnt["DMUL"]++
  if (line ~ /DADD/)   cnt["DADD"]++
  n++
}
function report() {
  if (cnt["DMMA"]+cnt["LDGSTS"]+cnt["UTMALDG"]+cnt["DFMA"] == 0) return
  printf "\n== %s\n   instructions %d | DMMA %d | LDGSTS %d | UTMALDG %d | SYNCS %d | LDS %d | STS %d | BAR.SYNC %d | DFMA %d | DMUL %d | DADD %d\n", name, n, cnt["DMMA"], cnt["LDGSTS"], cnt["UTMALDG"], cnt["SYNCS"], cnt["LDS"], cnt["STS"], cnt["BAR.SYNC"], cnt["DFMA"], cnt["DMUL"], cnt["DADD"]
  if (first_dmma != "") print "  " first_dmma
  if (first_cp != "")   print "  " first_cp
  if (first_tma != "")  print "  " first_tma
  if (first_sy != "")   print "  " first_sy
}
END { report() }'

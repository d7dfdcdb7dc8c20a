# usage: bv.sh name "flags"
cd /root/repo
DTFILL_NVCC_EXTRA="$2" python -c "
import sys
sys.path.insert(0,'.')
from distancetransform_depthcompletion_b200 import build as b
b.LIB_PATH='_variants/libdtfill_$1.so'
b.build(force=True, verbose=True)
" 2>&1 | grep -A2 "k2_chamferILi20ELb0ELb0ELb1E" | grep -E "Used|spill" 

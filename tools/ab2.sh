python tools/variant_ab.py "" 2>&1 | tail -1
ODL_WIDE_BLOCK=384 python tools/variant_ab.py "-DODL_BDF_THREADS=384" "-DODL_BDF_THREADS=384@tail_warps=30" "-DODL_BDF_THREADS=384@tail_warps=26" 2>&1 | tail -3
ODL_WIDE_BLOCK=512 python tools/variant_ab.py "-DODL_BDF_THREADS=512" "-DODL_BDF_THREADS=512@tail_warps=28" 2>&1 | tail -2
ODL_WIDE_BLOCK=128 python tools/variant_ab.py "@tail_warps=49" 2>&1 | tail -1

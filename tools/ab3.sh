for sms in 34 36 40 42 44; do echo "tail_sms $sms"; python tools/debug_beside.py tail_warps=$sms 2>&1 | tail -2; done

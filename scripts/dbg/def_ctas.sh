#!/bin/bash
for c in 4 5; do echo -n "ctas/SM $c, 1 GiB: "; DCMP_DEF_CTAS=$c python bench_extra.py --what c3 --c3-mib 1024 2>/dev/null | grep "^{" | cut -c60-200; done

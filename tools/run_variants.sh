#!/bin/bash
# time build/variants/lib_<name>.so for each name given (tools/time_all.py), one line per variant and stencil
for v in "$@"; do PARADIS_SL_LIB=build/variants/lib_$v.so python tools/time_all.py 2>&1 | sed "s/^/$v: /" | cut -c1-200; done

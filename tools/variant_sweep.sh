#!/bin/bash
# usage: bash tools/variant_sweep.sh [-c config] [variant ...]   ('' = the default library); prints the stage times of each
cfg=tnt-3m
if [ "$1" = "-c" ]; then cfg=$2; shift 2; fi
for v in "" "$@"; do
  if [ -z "$v" ]; then unset GS2M_LIB; else export GS2M_LIB=$PWD/gs-2m_b200/lib/variants/$v.so; fi
  echo "== variant '$v' ($cfg)"; python tools/stage_times.py $cfg 10 2>&1 | grep -v "^wall\|^torch\|^R ="
done

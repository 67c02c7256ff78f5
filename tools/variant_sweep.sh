#!/bin/bash
# usage: bash tools/variant_sweep.sh [variant ...]   ('' = the default library); prints the blend stage times of each
for v in "" "$@"; do
  if [ -z "$v" ]; then unset GS2M_LIB; else export GS2M_LIB=$PWD/gs-2m_b200/lib/variants/$v.so; fi
  echo "== variant '$v'"; python tools/stage_times.py tnt-3m 10 2>&1 | grep -i "blend_\|total"
done

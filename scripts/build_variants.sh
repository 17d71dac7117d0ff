#!/bin/bash
# Build kernel variants into variants/<name>.so for the same-box A/B (scripts/ab_variants.sh). Run in the build container.
#   scripts/build_variants.sh name1="-DFLAG=1 -DOTHER=2" name2="" ...
# A value of the form REV:<git rev> builds that revision in a scratch worktree (e.g. base=REV:HEAD).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/variants"
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  if [[ "$flags" == REV:* ]]; then
    rev="${flags#REV:}"; wt=/tmp/k9_variant_$name
    rm -rf "$wt"; git -C "$ROOT" worktree prune; git -C "$ROOT" worktree add -f "$wt" "$rev" >/dev/null 2>&1
    (cd "$wt" && python -m ka9q_sdr_b200.build >/dev/null)
    cp "$wt/ka9q_sdr_b200/libka9q_b200.so" "$ROOT/variants/$name.so"
    git -C "$ROOT" worktree remove --force "$wt"
  else
    tmp=/tmp/k9_variant_$name; rm -rf "$tmp"; mkdir -p "$tmp"
    cp -r "$ROOT/ka9q_sdr_b200" "$ROOT/include" "$tmp/"; rm -rf "$tmp/ka9q_sdr_b200/build" "$tmp/ka9q_sdr_b200/libka9q_b200.so"
    (cd "$tmp" && KA9Q_B200_NVCC_EXTRA="$flags" python -m ka9q_sdr_b200.build >/dev/null)
    cp "$tmp/ka9q_sdr_b200/libka9q_b200.so" "$ROOT/variants/$name.so"
    rm -rf "$tmp"
  fi
  echo "built variants/$name.so  [$flags]"
done

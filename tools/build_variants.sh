#!/bin/bash
# builds the tuning variants of tools/sweep.py from tools/variants.txt: one line per variant, "name defines..."
rm -rf kanpyo_b200/_variants
while read -r name defs; do
  [ -z "$name" ] && continue
  case "$name" in \#*) continue;; esac
  KP_VARIANT=$name KP_DEFINES="$defs" python -W ignore -m kanpyo_b200.build > /dev/null 2>&1 || echo "variant $name failed"
done < ${1:-tools/variants.txt}
ls kanpyo_b200/_variants

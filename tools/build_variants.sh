#!/bin/bash
# builds the tuning variants of tools/sweep.py: one line per variant, "name defines..."
rm -rf kanpyo_b200/_variants
while read -r name defs; do
  [ -z "$name" ] && continue
  KP_VARIANT=$name KP_DEFINES="$defs" python -m kanpyo_b200.build > /dev/null || echo "variant $name failed"
done <<'LIST'
v00_old -DKP_VIT_PRED=0 -DKP_WALK_ILP=0 -DKP_WALK_T1=0 -DKP_FILL_FLAT=0 -DKP_BK_PIPE=0 -DKP_CNT_MINB=8 -DKP_FILL_MINB=8 -DKP_BT_ORDER=0
v01_ilp -DKP_WALK_T1=0 -DKP_FILL_FLAT=0
v02_ilp_t1 -DKP_FILL_FLAT=0
v03_flat -DKP_WALK_ILP=0 -DKP_WALK_T1=0
v04_cnt8 -DKP_CNT_MINB=8
v05_cnt5 -DKP_CNT_MINB=5
v06_fill6 -DKP_FILL_MINB=6
v07_fill4 -DKP_FILL_MINB=4
v08_vg4 -DKP_VIT_GROUP=4
v09_vg16 -DKP_VIT_GROUP=16
v10_vm16 -DKP_VIT_MINB=16
v11_bt4 -DKP_BT_GROUP=4
v12_btnoord -DKP_BT_ORDER=0
LIST
ls -la kanpyo_b200/_variants

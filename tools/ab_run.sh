V=rl-selfplay-mnk_b200/build/variants
MNK_SAVE=/tmp/base.pt MNK_SIZES=2 python tools/time_tower.py 2>&1 | grep "envs=32768" | cut -c1-110
for n in "$@"; do echo "== $n"; MNK_LIB=$V/lib_$n.so MNK_CHECK=/tmp/base.pt MNK_SIZES=2 python tools/time_tower.py 2>&1 | grep "envs=32768\|variant\|Error\|error" | cut -c1-110; done

#!/bin/bash
# profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
#   UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync
# Run from the repo root after build():  bash scripts/sass_summary.sh > profiles/sass_summary.txt
SO=mlx-swift-qwen3-tts_b200/qwen3tts_b200/libqwen3tts_b200.so
echo "# $(basename $SO): cuobjdump -sass, mnemonic counts per kernel (sm_100a)"
echo "# kernel | UTC*MMA | LDTM | STTM | UTMALDG | UBLKCP | UTCBAR | HMMA | SYNCS | instructions"
cuobjdump -sass "$SO" | awk '
  /Function :/ { if (name != "") print name " | " mma " | " ldtm " | " sttm " | " tma " | " blk " | " bar " | " hmma " | " syncs " | " n;
                 name=$3; mma=0; ldtm=0; sttm=0; tma=0; blk=0; bar=0; hmma=0; syncs=0; n=0 }
  /^[ \t]+\/\*[0-9a-f]+\*\// { n++ }
  /UTC[A-Z]*MMA/ { mma++ } /LDTM/ { ldtm++ } /STTM/ { sttm++ } /UTMALDG/ { tma++ } /UBLKCP/ { blk++ } /UTCBAR/ { bar++ } / HMMA/ { hmma++ } /SYNCS/ { syncs++ }
  END { print name " | " mma " | " ldtm " | " sttm " | " tma " | " blk " | " bar " | " hmma " | " syncs " | " n }' | c++filt | sort -t'|' -k2 -nr

#!/bin/bash
# compute-sanitizer over a small batch case and a small fuzz seed.  usage: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import sys, os, random
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "tests")]
import numpy as np
import textgen, cases
from _oracle import Oracle
import wordpiece_b200
mode = sys.argv[1]
if mode == "batch":
    vocab = ["[UNK]", "a", "b", "ab", "##b", "##c", "abc", "self", "-", "made", "中"]
    v = wordpiece_b200.Vocab(vocab, device=0)
    texts = [b"self-made", b"ab abc", b"", b"a" * 700 + b" b", "中ab".encode()]
    ids, offs = v.encode_batch(texts)
    o = Oracle(vocab)
    ok = all(np.array_equal(o.encode(t), ids[int(offs[i]):int(offs[i + 1])]) for i, t in enumerate(texts))
    print("batch ok" if ok else "batch MISMATCH", ids.tolist()[:20], offs.tolist())
else:
    text, vocab = textgen.case(1140, 60000, invalid_rate=0.0, long_run_rate=0.05, long_tokens=5)
    os.environ["WORDPIECE_B200_PIPE_CHUNK"] = "4096"
    v = wordpiece_b200.Vocab(vocab, device=0)
    exp = Oracle(vocab).encode(text)
    out = np.full(len(exp) + 64, -7, np.int32)
    k = v.encode_into(text, out)
    print("pipe ok" if np.array_equal(exp, out[:k]) else f"pipe MISMATCH {len(exp)} vs {k}")
    got = v.encode(text)
    print("single ok" if np.array_equal(exp, got) else f"single MISMATCH {len(exp)} vs {len(got)}")
PY
for tool in initcheck memcheck racecheck; do
  for mode in batch pipe; do
    timeout -k 10 600 compute-sanitizer --tool $tool --print-limit 30 python /tmp/san_case.py $mode > $OUT/san_${tool}_${mode}_$TAG.log 2>&1
    echo "== $tool $mode rc=$? : $(grep -E 'ERROR SUMMARY|ok$|MISMATCH|ok ' $OUT/san_${tool}_${mode}_$TAG.log | tr '\n' ' ')"
    grep -E "Uninitialized|Invalid|hazard|at .*wp_|in .*wp_encode.cu" $OUT/san_${tool}_${mode}_$TAG.log | sed 's/^=* *//' | sort | uniq -c | sort -rn | head -12
  done
done

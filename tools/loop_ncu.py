"""One small loop check under ncu: python tools/loop_ncu.py K   (ncu --metrics gpu__time_duration.sum ...)"""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import datagen
import slam_loop_closing_b200 as lc
K = int(sys.argv[1]) if len(sys.argv) > 1 else 67
q = datagen.uniform(2000, 1234)
with lc.Context(loop_threshold=0.15, min_loop_gap=0) as c:
    c.db_append_many(list(range(K)), [datagen.uniform(2000, 5678 + k) for k in range(K)])
    for _ in range(4):
        c.loop_check(q, K + 30)

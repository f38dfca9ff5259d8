"""Indel tests from the UNMODIFIED reference (oracle/_ref/libsnpref.so): tests/golden/indel_tests.npz.
Run where /root/reference is mounted:  python tests/golden/make_golden_indel.py
Each case is one event test of call_indels (lofreq_call.c:618-726): the column is rebuilt with the reference's own
add_ins_sequence / add_del_sequence, its error probabilities come from plp_to_ins_errprobs / plp_to_del_errprobs
(snpcaller.c:501-623), the p-value from snpcaller() with (event count, 0, 0).  Stored in the layout
lfb200_indel_tests() takes: reads of the tested event last, 255 = quality not available."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import Oracle, VARCALL_USE_MQ  # noqa: E402

USE_SQ, USE_IDAQ = 4, 8        # defaults.h: VARCALL_USE_SQ, VARCALL_USE_IDAQ


def nb(x):
    """reference int (-1 = not available) -> plane byte"""
    x = np.asarray(x, np.int64)
    return np.where(x < 0, 255, np.minimum(x, 254)).astype(np.uint8)


def main():
    ref = Oracle("reference")
    rng = np.random.default_rng(20261017)
    sig = float(np.float32(0.01))
    iq, mq, aq, sq, off, cnt, bonf, flags, isdel, pv, names = [], [], [], [], [0], [], [], [], [], [], []
    spec = []
    for is_del in (0, 1):
        for flag in (VARCALL_USE_MQ | USE_IDAQ, VARCALL_USE_MQ | USE_IDAQ | USE_SQ, VARCALL_USE_MQ, 0, USE_IDAQ):
            for n_other, sizes in ((40, (3,)), (400, (12, 3)), (500, (1, 1, 1)), (900, (60, 5, 2)), (2500, (300, 40)), (30, (30, 2))):
                spec.append((is_del, flag, n_other, sizes))
    for ci, (is_del, flag, n_other, sizes) in enumerate(spec):
        other_q = rng.integers(20, 46, n_other)
        other_mq = rng.choice([0, 20, 40, 60, 60, 60], n_other)
        events = []
        for s in sizes:
            events.append((rng.integers(15, 46, s), rng.choice([-1, 10, 25, 40, 60], s), rng.choice([0, 30, 60, 60, 255], s),
                           rng.choice([-1, 20, 40], s)))
        for te in range(len(sizes)):
            b = int(rng.choice([1, 7, 300, 30000, 3000000]))
            p, count, n_ep = ref.indel_test(is_del, flag, sig, b, other_q, other_mq, events, te)
            assert count == sizes[te] and n_ep == n_other + sum(sizes)
            # layout of lfb200_indel_tests: non-indel reads, the other events, the tested event last
            order = [e for e in range(len(sizes)) if e != te] + [te]
            iq.append(np.concatenate([nb(other_q)] + [nb(events[e][0]) for e in order]))
            # non-indel reads: mq as it is (snpcaller.c:524-526 has no 255 -> -1 mapping; 10^-25.5 vanishes in the merge)
            mq.append(np.concatenate([nb(np.where(other_mq == 255, -1, other_mq))] +
                                     [nb(np.where(events[e][2] == 255, -1, events[e][2])) for e in order]))
            aq.append(np.concatenate([np.full(n_other, 255, np.uint8)] +
                                     [np.full(len(events[e][0]), 255, np.uint8) for e in order[:-1]] + [nb(events[te][1])]))
            sq.append(np.concatenate([np.full(n_other, 255, np.uint8)] + [nb(events[e][3]) for e in order]))
            off.append(off[-1] + len(iq[-1]))
            cnt.append(count); bonf.append(b); flags.append(flag); isdel.append(is_del)
            pv.append(np.frombuffer(np.array([p], np.longdouble).tobytes(), np.uint8))
            names.append("c%d_e%d" % (ci, te))
    np.savez_compressed(os.path.join(HERE, "indel_tests.npz"), names=np.array(names), iq=np.concatenate(iq), mq=np.concatenate(mq),
                        aq=np.concatenate(aq), sq=np.concatenate(sq), read_off=np.array(off, np.int64), event_count=np.array(cnt, np.int32),
                        bonf=np.array(bonf, np.int64), flag=np.array(flags, np.int32), is_del=np.array(isdel, np.int32),
                        sig=np.float64(sig), pvalue_ld=np.stack(pv))
    print("wrote %d indel tests, %d reads" % (len(names), off[-1]))


if __name__ == "__main__":
    main()

"""Log-record parsing and ingest (host side of the path; mirrors rl4rs/utils/datautil.py:8-69).

The reference re-parses every '@'-separated record on every ``reset`` (base.py:92-99 ->
slate.py:67-83 -> datautil.py:20-32) and pads/truncates per row on every ``step``
(datautil.py:34-69).  Here a log file is parsed ONCE into a structure of arrays
(:class:`rl4rs_b200.synth.LogSoA`) that is uploaded to HBM; the per-step feature assembly then
happens in the CUDA kernels (csrc/r4_kernels.cu: ``k_assemble``).
"""
import numpy as np

from ..synth import LogSoA


class FeatureUtil(object):
    """Same constructor and ``record_split`` contract as the reference's FeatureUtil."""

    def __init__(self, config):
        self.config = config
        self.maxlen = config.get("maxlen", 64)
        self.batch_size = config["batch_size"]
        self.class_num = config.get("class_num", 2)
        self.dense_feature_num = config.get("dense_feature_num", 432)
        self.category_feature_num = config.get("category_feature_num", 21)
        self.category_hash_size = config.get("category_hash_size", 100000)
        self.seq_num = config.get("seq_num", 2)

    @classmethod
    def record_split(cls, record):
        """One record -> 9 typed fields (datautil.py:20-32)."""
        f = record.split("@")
        if len(f) != 9:
            raise ValueError("log record must have 9 '@'-separated fields, got %d" % len(f))
        return (int(f[0]), int(f[1]), int(f[2]),
                [int(x) for x in f[3].split(",")],
                [int(x) for x in f[4].split(",")],
                [int(x) for x in f[5].split(",")],
                [float(x) for x in f[6].split(",")],
                [float(x) for x in f[7].replace(";", ",").split(",")],
                int(f[8]))

    def feature_extraction(self, data):
        """Raw 6-field state rows -> ((sequence i32 [n, seq_num, maxlen], dense f32 [n, dense_feature_num], category i32
        [n, category_feature_num], slate_label), labels) exactly like datautil.py:34-69: each behaviour sequence pre-padded
        with 0 / cut to its LAST maxlen ids, dense and category rows post-padded / post-truncated.  The env itself never
        calls this -- its feature rows are assembled on the GPU (k_assemble, k_seq_ids) -- it exists for `obs_fn` plug-ins
        written against the reference (slate.py:244-249 calls it on `state['state']`), e.g. over `samples.state`."""
        n = len(data)
        seqs = np.zeros((n, self.seq_num, self.maxlen), np.int32)
        dense = np.zeros((n, self.dense_feature_num), np.float32)
        cat = np.zeros((n, self.category_feature_num), np.int32)
        slate_labels, labels = [], []
        for i, (role_id, sequence_feature, dense_feature, category_feature, slate_label, label) in enumerate(data):
            for j, s in enumerate(list(sequence_feature)[:self.seq_num]):
                s = np.asarray(s, dtype=np.int64)[-self.maxlen:]
                if len(s):
                    seqs[i, j, self.maxlen - len(s):] = s
            d = np.asarray(dense_feature, dtype=np.float32)[:self.dense_feature_num]
            dense[i, :len(d)] = d
            c = np.asarray([int(x) for x in category_feature], dtype=np.int64)[:self.category_feature_num]
            cat[i, :len(c)] = c
            slate_labels.append(slate_label)
            labels.append(label)
        return (seqs, dense, cat, np.array(slate_labels)), labels

    @staticmethod
    def parse_log(lines, maxlen=64):
        """Text records -> LogSoA.  Only the fields the env reads are kept (SURVEY.md Appendix A:
        fields 1, 3, 4, 5, 6): session id, exposed items, feedback, history, portrait.

        History is pre-padded with 0 / truncated to the LAST ``maxlen`` ids exactly as
        ``pad_sequences(..., maxlen)`` does at datautil.py:43-46; portrait splits into 10
        categorical ids (parsed float -> int, slate.py:77 + datautil.py:49) and 32 dense floats
        (f64 text -> f32, datautil.py:52-58)."""
        lines = [ln.rstrip("\r\n") for ln in lines]
        # the reference's reader treats the FIRST empty line as end of file and rewinds (base.py:85-88): rows behind a
        # blank line are never served, so they are not ingested either
        for i, ln in enumerate(lines):
            if not ln.strip():
                lines = lines[:i]
                break
        n = len(lines)
        S = len(lines[0].split("@")[3].split(",")) if n else 9
        ts = np.zeros(n, np.int64)
        sess = np.zeros(n, np.int64)
        sid = np.zeros(n, np.int32)
        user_cat = np.zeros((n, 10), np.int32)
        user_dense = np.zeros((n, 32), np.float32)
        user_seq = np.zeros((n, maxlen), np.int32)
        seq_len = np.zeros(n, np.int32)
        items = np.zeros((n, S), np.int32)
        feedback = np.zeros((n, S), np.uint8)
        hist = []
        for i, ln in enumerate(lines):
            f = ln.split("@")
            if len(f) != 9:
                raise ValueError("line %d: expected 9 '@'-separated fields, got %d" % (i, len(f)))
            ts[i], sess[i], sid[i] = int(f[0]), int(f[1]), int(f[2])
            it = np.array(f[3].split(","), dtype=np.int64)
            fb = np.array(f[4].split(","), dtype=np.int64)
            if len(it) != S or len(fb) != S:
                raise ValueError("line %d: %d items / %d feedbacks, expected %d" % (i, len(it), len(fb), S))
            items[i], feedback[i] = it, fb
            h = np.array(f[5].split(","), dtype=np.int64)
            hist.append(h)
            seq_len[i] = len(h)
            h = h[-maxlen:]
            user_seq[i, maxlen - len(h):] = h
            p = np.array(f[6].split(","), dtype=np.float64)
            if len(p) != 42:
                raise ValueError("line %d: user_protrait has %d values, expected 42" % (i, len(p)))
            user_cat[i] = p[:10].astype(np.int64)
            user_dense[i] = p[10:].astype(np.float32)
        log = LogSoA(timestamp=ts, session_id=sess, sequence_id=sid, user_cat=user_cat,
                     user_dense=user_dense, user_seq=user_seq, seq_len=seq_len, items=items,
                     feedback=feedback, hist=hist)
        log.lines = list(lines)          # the record strings: samples.records / to_string() (base.py:28-31,56-57)
        return log

    @staticmethod
    def load_log(path, maxlen=64):
        """``sample_file``: a text log (reference format, no header) or a ``.npz`` LogSoA."""
        if str(path).endswith(".npz"):
            return LogSoA.load(path)
        with open(path, "r") as f:
            return FeatureUtil.parse_log(f.read().split("\n"), maxlen)

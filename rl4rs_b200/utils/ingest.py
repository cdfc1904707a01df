"""Log ingest before the path: the reference's session-level preprocessing, restated over
lists of record strings, plus the one-off text -> structure-of-arrays conversion the GPU loader reads.

Reference: ``script/data_preprocess.py``
  * ``slate2trajectory`` (:48-88)  joins the 4 page records of a session (dataset b2, one 9-item page per line) into ONE
    trajectory record with 36 exposed items / feedbacks and ``;``-joined item features (dataset b3) -- the rows
    ``SeqSlateRecEnv`` replays.  Quirks kept: header lines (containing 'timestamp') and empty lines are skipped; a
    session is flushed only when the NEXT session starts, so the file's LAST session is never written; a session that
    does not hold exactly 4 records is an assertion error.
  * ``data_augment`` (:6-45)  pads sessions with fewer than 4 pages: the missing pages take the exposed items and item
    features of a random other line (``np.random.randint(1, data_size - 1)`` on the GLOBAL numpy RNG), zero feedback,
    timestamp + 1 and sequence_id + 1 of the session's last page.  Same flush-on-next-session quirk.
The reference re-reads and re-splits the text on every ``reset`` (base.py:82-108, datautil.py:20-32); here a file is
parsed once (``ingest``) into the ``.npz`` LogSoA that ``config['sample_file']`` accepts.
"""
import numpy as np

from .datautil import FeatureUtil


def _records(lines):
    for record in lines:
        if len(record) < 1 or "timestamp" in record:              # data_preprocess.py:14-15,56-57
            continue
        yield record


def slate2trajectory(lines):
    """[page records, session-contiguous] -> [trajectory records] (data_preprocess.py:48-88)."""
    out, tmp, prev = [], [], None
    for record in _records(lines):
        role_id = record.split("@")[1]
        if role_id == prev or prev is None:
            tmp.append(record)
            prev = role_id
            continue
        if len(tmp) != 4:
            raise AssertionError("session %s has %d page records, slate2trajectory needs exactly 4" % (prev, len(tmp)))
        f = [x.split("@") for x in tmp]
        out.append("@".join([f[0][0], f[0][1], "1", ",".join(x[3] for x in f), ",".join(x[4] for x in f),
                             f[0][5], f[0][6], ";".join(x[7] for x in f), f[0][8]]))
        tmp, prev = [record], role_id
    return out                                                     # the last session is NOT flushed (reference quirk)


def data_augment(lines):
    """Pad every session to 4 pages (data_preprocess.py:6-45); consumes the global numpy RNG like the reference."""
    data = list(lines)
    data_size = len(data)
    out, tmp, prev = [], [], None
    for record in data:
        if len(record) < 1 or "timestamp" in record:
            continue
        role_id = record.split("@")[1]
        if role_id == prev or prev is None:
            tmp.append(record)
            prev = role_id
            continue
        assert len(tmp) <= 4
        for _ in range(len(tmp), 4):
            ts, sess, sid, _, _, useq, portrait, _, pol = tmp[-1].split("@")
            j = np.random.randint(1, data_size - 1)
            other = data[j].split("@")
            tmp.append("@".join([str(int(ts) + 1), sess, str(int(sid) + 1), other[3], "0,0,0,0,0,0,0,0,0",
                                 useq, portrait, other[7], pol]))
        out.extend(tmp)
        tmp, prev = [record], role_id
    return out


def ingest(lines_or_path, out_npz=None, maxlen=64, trajectories=False):
    """Text log (file path or list of lines) -> LogSoA; ``trajectories=True`` applies ``slate2trajectory`` first.
    With ``out_npz`` the arrays are saved in the format ``config['sample_file'] = '....npz'`` loads."""
    if isinstance(lines_or_path, str):
        with open(lines_or_path, "r") as f:
            lines = f.read().split("\n")
    else:
        lines = list(lines_or_path)
    if trajectories:
        lines = slate2trajectory(lines)
    else:
        lines = [ln for ln in lines if "timestamp" not in ln]       # run_split.sh strips the header the same way
    log = FeatureUtil.parse_log(lines, maxlen)
    if out_npz:
        log.save(out_npz)
    return log

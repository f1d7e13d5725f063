"""CPU oracle for the frame-window / loader rows (L0-L2).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in numpy,
  src/utils/ibl_data_utils.py:958-967   per-trial frame indices: reg_frame_num = int(fps * interval_len),
                                        start = np.searchsorted(ts, trial[0]), arange(start, start + reg_frame_num),
                                        ValueError when the trial holds > 10 frames too many / too few
  src/loader/base.py:39,50-63           channel 0 of the decoded (T,H,W,C) uint8 video -> (T,1,H,W) -> .float()
  src/trainer/base.py:64-67             cat([batch[mod].flatten(1) ...], -1)
The reference has no tests for these; parity is pinned by construction (the functions are literal
transcriptions) and by the golden sorted_idx hash for the frame subset (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np


def load_video_index(ts: np.ndarray, intervals: np.ndarray, fps: float) -> np.ndarray:
    interval_len = intervals[0, 1] - intervals[0, 0]
    reg_frame_num = int(fps * interval_len)
    trial_index_list = []
    for trial in intervals:
        ts_trial = ts[(ts > trial[0]) & (ts < trial[1])]
        start_idx = np.searchsorted(ts, trial[0])
        frame_idxes = np.arange(start_idx, start_idx + reg_frame_num)
        if abs(len(ts_trial) - reg_frame_num) > 10:
            raise ValueError("Number of frames in the video does not match the expected number of frames")
        trial_index_list.append(frame_idxes)
    return np.array(trial_index_list)


def cut_windows(session_frames: np.ndarray, trial_index_list: np.ndarray) -> np.ndarray:
    """(n_frames, H, W) uint8 -> (n_trials, 120, H, W) uint8 by plain fancy indexing."""
    return session_frames[trial_index_list]


def loader_cast(video_THWC_u8: np.ndarray) -> np.ndarray:
    """src/loader/base.py:50-63 + :39: first channel, (T,1,H,W), float32 without scaling."""
    return video_THWC_u8[:, :, :, 0][:, None].astype(np.float32)

"""``python -m lidal_b200.score.cli --dataset_name SK --model_name SPVCNN --r_id 1`` -- the reference's
``python -m score.sv_level.LiDAL`` (README.md:115; score/sv_level/LiDAL.py:104-330) on the CUDA scorer.

Same arguments, same ``Processing_files/<dataset>/...`` tree in and out (the file formats are the ABI between the
reference's steps, SURVEY.md section 8b, boundary B2):

  in   sv_flag/KMeans/0r/<seq>/*.npy                                     (r_id == 1)          LiDAL.py:152-155
       sv_flag/KMeans/<model>/LiDAL/<r_id-1>r/<seq>/*.npy                (r_id  > 1)
       prob_map/<model>/fr/0r/<seq>/*.npy | prob_map/<model>/sv/LiDAL/<r_id-1>r/<seq>/*.npy   LiDAL.py:187-191
       kdtree/<seq>/*.pickle, super_voxel/KMeans/<seq>/*.pickle                               LiDAL.py:193-197
       super_voxel/KMeans/sv_pnums.npy, sv_centers.npy  (cache; written when absent)          LiDAL.py:171-180,220-222
  out  sv_flag/KMeans/<model>/LiDAL/<r_id>r/<seq>/<frame>.npy  (int flags 0 / 1 / 2)          LiDAL.py:326-330

What differs from the reference: one GPU replaces the 24-process pool (every sequence becomes resident in HBM, is scored
and dropped), the two greedy passes run through ``select_regions`` (device sorts and 5 m pair lists, native host walk),
and nothing is printed per candidate.  ``run`` takes the scorer and the selector as arguments so that the orchestration
and the file handling are testable on a CPU against the reference's own ``__main__`` (tests/test_cli_cpu.py).
"""
from __future__ import annotations

import argparse
import glob
import os

import numpy as np

SK_TRAIN_SPLIT = ['00', '01', '02', '03', '04', '05', '06', '07', '09', '10']          # LiDAL.py:124
TRAIN_POINT_NUM = {"SK": 2349559532, "NU": 976677792}                                     # LiDAL.py:126,131
NEI_NUM, DIS_THRESH = 24, 0.1                                                             # LiDAL.py:117-120


def train_split_of(dataset_name: str):
    if dataset_name == "SK":
        return list(SK_TRAIN_SPLIT)
    if dataset_name == "NU":
        from nuscenes.utils.splits import create_splits_scenes                            # LiDAL.py:129-130
        return create_splits_scenes()["train"]
    raise ValueError(f"unknown dataset {dataset_name!r} (the reference knows 'SK' and 'NU')")


def run(dataset_name: str, model_name: str, r_id: int, root: str = ".", train_split=None, train_point_num=None,
        score_dataset=None, select_regions=None, device="cuda", verbose=True):
    """LiDAL.py:116-330.  Returns (sv_flags int [R], paths written).  ``score_dataset(sequences, nei_num, dis_thresh,
    n_regions_total, sv_pnums, sv_centers)`` and ``select_regions(sv_flags, D, H, pnums, centres, train_point_num)`` default
    to the CUDA implementations of this package."""
    assert r_id >= 1                                                                       # LiDAL.py:150
    if score_dataset is None or select_regions is None:
        from . import score_dataset as _score_dataset, select_regions as _select_regions
        score_dataset = score_dataset or (lambda *a, **k: _score_dataset(*a, device=device, **k))
        select_regions = select_regions or (lambda *a, **k: _select_regions(*a, device=device, **k))
    train_split = train_split_of(dataset_name) if train_split is None else list(train_split)
    train_point_num = TRAIN_POINT_NUM[dataset_name] if train_point_num is None else train_point_num
    base = os.path.join(root, "Processing_files", dataset_name)

    # ---- current labelled set: one flag file per frame, concatenated in train_split / file-name order   (:137-168)
    sv_flags, frame_sv_offsets, save_paths = [], [0], []
    for seq_id in train_split:
        if r_id == 1:
            flag_files = sorted(glob.glob(os.path.join(base, "sv_flag", "KMeans", "0r", seq_id, "*.npy")))
        else:
            flag_files = sorted(glob.glob(os.path.join(base, "sv_flag", "KMeans", model_name, "LiDAL", f"{r_id - 1}r", seq_id, "*.npy")))
        save_folder = os.path.join(base, "sv_flag", "KMeans", model_name, "LiDAL", f"{r_id}r", seq_id)
        os.makedirs(save_folder, exist_ok=True)
        for f_file in flag_files:
            sv_f = np.load(f_file)
            sv_flags.append(sv_f)
            frame_sv_offsets.append(frame_sv_offsets[-1] + sv_f.shape[0])
            save_paths.append(os.path.join(save_folder, os.path.basename(f_file)))
    sv_flags = np.concatenate(sv_flags).astype(np.float64) if sv_flags else np.zeros(0)    # np.append's float64 (:165)
    n_regions = sv_flags.shape[0]

    # ---- cached region statistics                                                                   (:171-180)
    pn_path = os.path.join(base, "super_voxel", "KMeans", "sv_pnums.npy")
    c_path = os.path.join(base, "super_voxel", "KMeans", "sv_centers.npy")
    sv_pnums = sv_centers = None
    if os.path.exists(pn_path):
        sv_pnums, sv_centers = np.load(pn_path), np.load(c_path)

    # ---- per-sequence file lists                                                                    (:185-200)
    sequences = []
    for seq_id in train_split:
        if r_id == 1:
            prob_files = sorted(glob.glob(os.path.join(base, "prob_map", model_name, "fr", "0r", seq_id, "*.npy")))
        else:
            prob_files = sorted(glob.glob(os.path.join(base, "prob_map", model_name, "sv", "LiDAL", f"{r_id - 1}r", seq_id, "*.npy")))
        kdtree_files = sorted(glob.glob(os.path.join(base, "kdtree", seq_id, "*.pickle")))
        sv_info_files = sorted(glob.glob(os.path.join(base, "super_voxel", "KMeans", seq_id, "*.pickle")))
        assert len(prob_files) == len(kdtree_files)
        assert len(kdtree_files) == len(sv_info_files)
        if verbose:
            print(f"sequence {seq_id}: {len(prob_files)} frames")
        sequences.append((prob_files, kdtree_files, sv_info_files))

    # ---- inter-frame divergence / entropy of every region                                           (:202-218)
    sv_interds, sv_interes, pn_out, c_out, sv_pre = score_dataset(sequences, NEI_NUM, DIS_THRESH, n_regions, sv_pnums, sv_centers)
    if not sv_pre:                                                                                    # :220-222
        sv_pnums, sv_centers = pn_out, c_out
        np.save(pn_path, sv_pnums)
        np.save(c_path, sv_centers)

    # ---- the two greedy passes                                                                      (:230-325)
    flags = select_regions(sv_flags, sv_interds, sv_interes, sv_pnums, sv_centers, train_point_num)

    # ---- one flag file per frame, named like its input                                              (:327-330)
    for idx in range(len(frame_sv_offsets) - 1):
        np.save(save_paths[idx], flags[frame_sv_offsets[idx]:frame_sv_offsets[idx + 1]])
    if verbose:
        print(f"{n_regions} regions: {int((flags == 1).sum())} labelled, {int((flags == 2).sum())} pseudo-labelled; "
              f"{len(save_paths)} flag files written")
    return flags, save_paths


def main(argv=None):
    ap = argparse.ArgumentParser(description="Active selection with pseudo labels in the super voxel level based on inter-scan "
                                             "divergence and inter-scan entropy (B200 scorer)")
    ap.add_argument("--dataset_name", type=str, required=True, help="name of the used dataset")
    ap.add_argument("--model_name", type=str, required=True, help="name of the trained model providing prob inference")
    ap.add_argument("--r_id", type=int, required=True, help="current training r_id")
    ap.add_argument("--root", type=str, default=".", help="directory that holds Processing_files/ (the reference uses the cwd)")
    args = ap.parse_args(argv)
    run(args.dataset_name, args.model_name, args.r_id, root=args.root)


if __name__ == "__main__":
    main()

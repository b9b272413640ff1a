"""Stand-in for the absent nuscenes-devkit so score/sv_level/LiDAL.py:10 imports (oracle pinning only)."""


def create_splits_scenes():
    return {"train": [], "val": []}

"""Offline bulk encoder: the host-side file logic -- chunking, file names, the split over the ranks, the fps
thinning and the trainer's read -- against the reference's own expressions (videollama2_arch.py:236-276,
process_clip_encoder.py:49-57, videollama2_trainer_score.py:312-315), with the CUDA tower replaced by a stand-in."""
import os

import torch

from streammind_b200 import offline


def test_rank_split_is_the_reference_expression():
    vids = [f"v{i}" for i in range(11)]
    for world in (1, 2, 3, 4, 8):
        local_batch = len(vids) // world                                   # videollama2_arch.py:236
        for rank in range(world):
            assert offline.rank_slice(vids, rank, world) == vids[rank * local_batch:(rank + 1) * local_batch]   # :239


def test_chunks_and_names():
    assert offline.chunk_ranges(1234) == [(0, 500, 500), (500, 1000, 1000), (1000, 1234, 1500)]
    assert offline.chunk_ranges(500) == [(0, 500, 500)]
    assert offline.chunk_ranges(2100, 1000) == [(0, 1000, 1000), (1000, 2000, 2000), (2000, 2100, 3000)]
    vp = "/mnt/input/MatchTime/features_video/england_epl/2015-02-21_Chelsea/1_224p.mkv"
    half = os.path.basename(vp).split("_224p.mkv")[0]                      # :241-242
    assert offline.half_of(vp) == half == "1"
    d = os.path.dirname(vp.replace("features_video", "features_video_encode_ddp"))      # :271-272
    assert offline.encoded_dir(vp) == d
    assert offline.feature_file_name(half, 1000, 1500) == "{}_encode_feature_frame_{}_{}.pt".format(half, 1000, 1000 + 500)   # :274


def test_files_thinning_and_trainer_read(tmp_path, monkeypatch):
    T, P, C = 23, 4, 8
    feats_all = torch.arange(T * P * C, dtype=torch.float32).view(T, P, C)

    def fake_encode(engine, pixels, pooled=False):                          # "features" of frame i = row i of feats_all
        idx = pixels[:, 0, 0, 0].long()
        return feats_all[idx].mean(1) if pooled else feats_all[idx]

    monkeypatch.setattr(offline, "_encode", fake_encode)
    read = lambda ids: torch.tensor(ids, dtype=torch.float32).view(-1, 1, 1, 1).expand(-1, 3, 2, 2)
    root = tmp_path / "features_video_encode_ddp" / "game"
    paths = offline.encode_video_to_files(None, read, T, str(root), "1", chunk_frames=10)
    assert [os.path.basename(p) for p in paths] == ["1_encode_feature_frame_0_10.pt", "1_encode_feature_frame_10_20.pt", "1_encode_feature_frame_20_30.pt"]
    full = [torch.load(p) for p in paths]
    assert [tuple(f.shape) for f in full] == [(1, 10, P, C), (1, 10, P, C), (1, 3, P, C)]
    assert torch.equal(torch.cat(full, 1)[0], feats_all)
    # thinning a saved file (process_clip_encoder.py) == encoding with segment
    thin = [torch.load(offline.thin_feature_file(p, 3)) for p in paths]
    direct = [torch.load(p) for p in offline.encode_video_to_files(None, read, T, str(tmp_path / "direct"), "1", chunk_frames=10, segment=3)]
    for a, b, f in zip(thin, direct, full):
        assert torch.equal(a, f[:, ::3]) and torch.equal(a, b)
    assert "features_video_encode_ddp_fps" in offline.thin_feature_file(paths[0], 3)
    # the trainer's read
    assert torch.equal(offline.load_feature_slice(paths[0], 2, 9, 3), full[0][:, 2:9:3])
    # pooled files
    pp = offline.encode_video_to_files(None, read, T, str(tmp_path / "pooled"), "2", chunk_frames=10, pooled=True)
    assert os.path.basename(pp[0]) == "2_encode_feature_frame_0_10_pooled.pt"
    assert torch.equal(torch.load(pp[2])[0], feats_all[20:].mean(1))


def test_all_videos_job_shards_by_file(tmp_path, monkeypatch):
    monkeypatch.setattr(offline, "_encode", lambda e, px, pooled=False: torch.zeros(px.shape[0], 2, 2))
    vids = [str(tmp_path / "features_video" / f"g{i}" / "1_224p.mkv") for i in range(5)]
    opened = []

    def open_video(p):
        opened.append(p)
        return (lambda ids: torch.zeros(len(ids), 3, 2, 2)), 7

    w0 = offline.encode_all_videos(None, vids, open_video, rank=0, world_size=2, chunk_frames=5)
    w1 = offline.encode_all_videos(None, vids, open_video, rank=1, world_size=2, chunk_frames=5)
    assert opened == vids[:4]                                               # 5 // 2 = 2 per rank, the fifth is nobody's (reference behaviour)
    assert len(w0) == len(w1) == 4 and all("features_video_encode_ddp" in p for p in w0 + w1)

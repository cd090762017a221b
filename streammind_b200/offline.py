"""Offline bulk encoder + feature files: the data-parallel job that turns whole videos into the
CLIP feature files the score trainer reads.  Mirrors, file name for file name and tensor for tensor,

  * ``encode_all_videos_score``           /root/reference/streammind/model/videollama2_arch.py:212-282
      500-frame chunks -> vision tower -> ``[1, t, 576, 1024]`` -> ``{half}_encode_feature_frame_{start}_{start+500}.pt``,
      videos split over the ranks as ``list[rank * (len // world) : (rank + 1) * (len // world)]``
  * the 1000-frame variant                /root/reference/streammind/encode_video_ori.py:544-591
  * the fps thinning ``[:, ::segment]``   /root/reference/process_clip_encoder.py:55-57,69-76 (segment = video_fps // fps = 12)
  * the trainer's read                    /root/reference/streammind/videollama2_trainer_score.py:312-315
      ``torch.load(path, map_location)[:, start:end:segment]``

Every frame goes through the same CUDA vision tower as the streaming path (``Engine.vit_encode``; batches of
``EngineConfig.max_frames`` frames per launch chain, no PyTorch arithmetic).  What differs by design:

  * ``segment``: the reference encodes all frames and thins the saved tensor afterwards; a frame's features do not depend on
    its neighbours, so encoding only frames ``start::segment`` writes the SAME ``*_fps`` tensor with 1/segment of the work
    (checked bit for bit in tests/test_offline_gpu.py);
  * ``pooled=True`` writes the mean over the 576 patches, ``[1, t, 1024]`` -- all the Mamba projector consumes (builder.py:
    403-414 starts with ``mean(dim=2)``) -- 576x smaller files; not a reference format, named ``*_pooled.pt``.

Decoding the video (decord) and the PIL preprocessing are the caller's: frames arrive as normalised pixels [T, 3, H, W] in the
model dtype, or as uint8 HWC frames that ``Engine.preprocess_frames`` (sm_preprocess_frames) turns into pixels on the device.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch

from .engine import Engine

CHUNK_FRAMES = 500          # videollama2_arch.py:247 (encode_video_ori.py uses 1000: pass chunk_frames=1000)


def rank_slice(items: Sequence, rank: int, world_size: int) -> List:
    """The reference's split of the video list over the ranks (videollama2_arch.py:236,239): equal shares of
    ``len // world_size``; the remainder at the end of the list is NOT encoded by anyone (as in the reference)."""
    local = len(items) // world_size
    return list(items[rank * local:(rank + 1) * local])


def chunk_ranges(duration: int, chunk_frames: int = CHUNK_FRAMES) -> List[Tuple[int, int, int]]:
    """(start, end of the frames actually present, end used in the FILE NAME) of every chunk: the name always says
    start + chunk_frames, the last chunk holds the frames up to the duration (videollama2_arch.py:247-250,274)."""
    return [(s, min(s + chunk_frames, duration), s + chunk_frames) for s in range(0, duration, chunk_frames)]


def feature_file_name(half: str, start: int, name_end: int, suffix: str = "") -> str:
    """``{half}_encode_feature_frame_{start}_{start+500}.pt`` (videollama2_arch.py:274; read back by
    videollama2_trainer_score.py:487,494); ``half`` is the part of the video file name before ``_224p.mkv`` (:242)."""
    return "{}_encode_feature_frame_{}_{}{}.pt".format(half, start, name_end, suffix)


def encoded_dir(video_path: str, src: str = "features_video", dst: str = "features_video_encode_ddp") -> str:
    """Directory of a video's feature files: the video's own directory with ``features_video`` replaced (:271-272)."""
    return os.path.dirname(video_path.replace(src, dst))


def half_of(video_path: str) -> str:
    return os.path.basename(video_path).split("_224p.mkv")[0]


def encode_frames(engine: Engine, pixels: torch.Tensor, pooled: bool = False) -> torch.Tensor:
    """[t, 3, H, W] pixels (device or pinned host, model dtype) -> [t, 576, 1024] patch features (or [t, 1024] pooled) on the
    device: ``CLIPVisionTower.forward`` over a whole chunk, ``max_frames`` frames per call of the CUDA tower."""
    c = engine.cfg
    t = pixels.shape[0]
    out = torch.empty((t, c.vit_hidden) if pooled else (t, c.num_patches, c.vit_hidden), dtype=c.dtype, device=engine.device)
    for i in range(0, t, c.max_frames):
        px = pixels[i:i + c.max_frames]
        if px.dtype != c.dtype:
            px = px.to(c.dtype)
        if not px.is_cuda:
            px = px.to(engine.device, non_blocking=True)
        feats, pl = engine.vit_encode(px.contiguous(), want_feats=not pooled)
        out[i:i + px.shape[0]] = pl if pooled else feats
    return out


_encode = encode_frames          # seam for the CPU tests of the file logic


def encode_video_to_files(engine: Engine, read_frames: Callable[[List[int]], torch.Tensor], duration: int, out_dir: str, half: str,
                          chunk_frames: int = CHUNK_FRAMES, segment: int = 1, pooled: bool = False, save=torch.save) -> List[str]:
    """One video -> its feature files.  ``read_frames(frame_id_list)`` returns the normalised pixels [n, 3, H, W] of those
    frames (the reference's decord + expand2square + CLIP preprocess, videollama2_arch.py:252-257).  ``segment > 1`` writes the
    fps-thinned tensor ``[:, ::segment]`` of every chunk directly.  Returns the paths written."""
    os.makedirs(out_dir, exist_ok=True)
    suffix = "_pooled" if pooled else ""
    paths = []
    for start, end, name_end in chunk_ranges(duration, chunk_frames):
        ids = list(range(start, end, segment))
        feats = _encode(engine, read_frames(ids), pooled=pooled).unsqueeze(0)        # [1, t, 576, 1024]: 'b t n h' with b = 1 (:266)
        path = os.path.join(out_dir, feature_file_name(half, start, name_end, suffix))
        if feats.is_cuda:
            torch.cuda.current_stream().synchronize()
        save(feats, path)
        paths.append(path)
    return paths


def thin_feature_file(path: str, segment: int, src: str = "features_video_encode_ddp", dst: str = "features_video_encode_ddp_fps",
                      map_location="cpu") -> str:
    """process_clip_encoder.py:49-57: load a feature file, keep every ``segment``-th frame, save under the ``_fps`` tree."""
    out = path.replace(src, dst)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    torch.save(torch.load(path, map_location=map_location)[:, ::segment].clone(), out)
    return out


def load_feature_slice(path: str, start_idx: Optional[int] = None, end_idx: Optional[int] = None, segment: int = 1, map_location="cpu"):
    """The trainer's read (videollama2_trainer_score.py:312-315)."""
    return torch.load(path, map_location=map_location)[:, start_idx:end_idx:segment]


def encode_all_videos(engine: Engine, video_paths: Sequence[str], open_video: Callable[[str], Tuple[Callable[[List[int]], torch.Tensor], int]],
                      rank: int = 0, world_size: int = 1, **kw) -> List[str]:
    """``encode_all_videos_score``: this rank's share of the videos, every one through ``encode_video_to_files``.
    ``open_video(path) -> (read_frames, duration)``.  No collective: the job shards by file."""
    written: List[str] = []
    for vp in rank_slice(video_paths, rank, world_size):
        read_frames, duration = open_video(vp)
        written += encode_video_to_files(engine, read_frames, duration, encoded_dir(vp), half_of(vp), **kw)
    return written

"""Host-side helpers of the streaming call, same names and argument meaning as the reference's
/root/reference/streammind/mm_utils.py (tokenizer_MMODAL_token :567-604, KeywordsStoppingCriteria
:616-647, process_video / process_image with aspect_ratio 'pad' :446-464, expand2square :257-268)."""
from __future__ import annotations

from typing import List, Sequence

import torch

from .constants import IMAGE_TOKEN_INDEX, MMODAL_INDEX_TOKEN


def tokenizer_MMODAL_token(prompt: str, tokenizer, MMODAL_token_index: int = IMAGE_TOKEN_INDEX, return_tensors=None):
    """Split ``prompt`` on the modality tag (``<video>`` for -201), tokenize each piece and join the id
    lists with the sentinel; a BOS that every piece starts with is kept once (mm_utils.py:590-598)."""
    tag = f"<{MMODAL_INDEX_TOKEN[MMODAL_token_index].lower()}>"
    pieces = [tokenizer(chunk).input_ids for chunk in prompt.split(tag)]
    ids: List[int] = []
    skip = 0
    if pieces and pieces[0] and pieces[0][0] == tokenizer.bos_token_id:
        skip = 1
        ids.append(pieces[0][0])
    for n, piece in enumerate(pieces):
        if n > 0:
            ids.append(MMODAL_token_index)
        ids.extend(piece[skip:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


class KeywordsStoppingCriteria:
    """Stop when the tail of the output equals a keyword's ids, or the decoded tail contains the
    keyword (mm_utils.py:631-641).  ``single_token_ids`` are the keywords the device-side greedy loop
    can test by itself; anything else is checked on the host by calling the object."""

    def __init__(self, keywords: Sequence[str], tokenizer, input_ids: torch.Tensor):
        self.keywords = list(keywords)
        self.tokenizer = tokenizer
        self.keyword_ids = []
        self.max_keyword_len = 0
        for kw in self.keywords:
            ids = list(tokenizer(kw).input_ids)
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.start_len = input_ids.shape[1]

    @property
    def single_token_ids(self) -> List[int]:
        return [int(k[0]) for k in self.keyword_ids if k.numel() == 1]

    @property
    def needs_host_check(self) -> bool:
        return any(k.numel() != 1 for k in self.keyword_ids)

    def call_for_batch(self, output_ids: torch.Tensor, scores=None, **kw) -> bool:
        offset = min(output_ids.shape[1] - self.start_len, self.max_keyword_len)
        for kid in self.keyword_ids:
            kid = kid.to(output_ids.device)
            if output_ids.shape[1] >= kid.shape[0] and bool((output_ids[0, -kid.shape[0]:] == kid).all()):
                return True
        if offset > 0:
            text = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
            return any(k in text for k in self.keywords)
        return False

    def __call__(self, output_ids: torch.Tensor, scores=None, **kw) -> bool:
        return all(self.call_for_batch(output_ids[i].unsqueeze(0), scores) for i in range(output_ids.shape[0]))


def process_video(frames, processor=None, aspect_ratio: str = "pad", *, engine) -> torch.Tensor:
    """mm_utils.process_video (:446-464) for frames that are already decoded: uint8 RGB [n, H, W, 3] (numpy array, list
    of arrays, CPU or CUDA tensor) -> pixel_values [n, 3, 336, 336] in the tower's dtype, on the engine's device.
    The reference pads every frame to a square of the processor's mean colour (expand2square), runs
    CLIPImageProcessor.preprocess on PIL images and casts to half; here the whole chain is one C-ABI call
    (sm_preprocess_frames, bit-exact with PIL 8-bit bicubic + the processor's float32 arithmetic).
    processor: anything with image_mean / image_std (e.g. the reference's CLIPImageProcessor); None = CLIP defaults."""
    if aspect_ratio != "pad":
        raise NotImplementedError("only aspect_ratio='pad' (the streaming demo's setting) is on the device path")
    import numpy as np
    if isinstance(frames, (list, tuple)):
        frames = np.stack([np.asarray(f) for f in frames])
    from .engine import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD
    mean = tuple(processor.image_mean) if processor is not None else OPENAI_CLIP_MEAN
    std = tuple(processor.image_std) if processor is not None else OPENAI_CLIP_STD
    return engine.preprocess_frames(frames, mean, std)


def process_image(image, processor=None, aspect_ratio: str = "pad", *, engine) -> torch.Tensor:
    """mm_utils.process_image: one RGB image [H, W, 3] -> [1, 3, 336, 336]."""
    import numpy as np
    return process_video(np.asarray(image)[None], processor, aspect_ratio, engine=engine)

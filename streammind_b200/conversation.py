"""Prompt template of the streaming demo: the reference's ``conv_templates['mistral_instruct']``
(/root/reference/streammind/conversation.py:383-393) rendered with the LLAMA_2 style (:78-98), which
hard-codes the instruction sentence at :90.  Only what the per-frame path needs is mirrored."""
from __future__ import annotations

import dataclasses
from enum import Enum, auto
from typing import List, Optional

INSTRUCTION = "Please describe the video content in detail based on the provided information."


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()
    LLAMA_2 = auto()


@dataclasses.dataclass
class Conversation:
    system: str
    roles: tuple
    messages: List[List[Optional[str]]]
    offset: int = 0
    sep_style: SeparatorStyle = SeparatorStyle.LLAMA_2
    sep: str = ""
    sep2: str = "</s>"
    version: str = "llama_v2"

    def copy(self) -> "Conversation":
        return dataclasses.replace(self, messages=[list(m) for m in self.messages])

    def append_message(self, role: str, message: Optional[str]):
        self.messages.append([role, message])

    def get_prompt(self) -> str:
        if self.sep_style != SeparatorStyle.LLAMA_2:
            raise NotImplementedError("only the LLAMA_2 style is on the streaming path")
        out = ""
        for i, (role, msg) in enumerate(self.messages):
            if i == 0:
                if not msg:
                    raise ValueError("first message should not be none")
                if role != self.roles[0]:
                    raise ValueError("first message should come from user")
            if not msg:
                continue
            if i == 0:
                msg = f"<<SYS>>\n{self.system}\n<</SYS>>\n\n" + INSTRUCTION + msg
            if i % 2 == 0:
                out += self.sep + f"[INST] {msg} [/INST]"
            else:
                out += " " + msg + " " + self.sep2
        return out.lstrip(self.sep) if self.sep else out


conv_mistral_instruct = Conversation(
    system="A chat between a curious user and an artificial intelligence assistant. "
           "The assistant gives helpful, detailed, and polite answers to the user's questions.",
    roles=("USER", "ASSISTANT"), messages=[])

conv_templates = {"mistral_instruct": conv_mistral_instruct, "llama_2": conv_mistral_instruct}

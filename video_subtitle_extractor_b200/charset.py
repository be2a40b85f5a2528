"""Class-id -> character tables (CTCLabelDecode: ['blank'] + dict lines + [' ']).

The dictionaries ship inside the paddleocr wheel, not in the reference tree.  ``en_dict.txt`` is reconstructible (95
printable ASCII characters, validated by decoding the reference's sample video); other languages need the dictionary
file (``rec_char_dict_path``) — without it the engine still returns class ids and the shim renders them as U+E000-based
private-use code points so that equality/similarity tests between frames (reference backend/main.py:949) keep working.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

EN_DICT = "".join(chr(c) for c in range(0x30, 0x7F)) + "".join(chr(c) for c in range(0x21, 0x30)) + " "


def characters(lang: str = "en", dict_path: Optional[str] = None, n_classes: Optional[int] = None) -> List[str]:
    if dict_path:
        with open(dict_path, "rb") as f:
            lines = [ln.decode("utf-8").strip("\n").strip("\r\n") for ln in f.readlines()]
        return ["blank"] + lines + [" "]
    if lang == "en" or n_classes == 97:
        return ["blank"] + list(EN_DICT[:-1]) + [" "] + [" "]
    n = n_classes or 0x1900
    return ["blank"] + [chr(0xE000 + i) for i in range(1, n)]


def ids_to_text(ids: Sequence[int], chars: Sequence[str]) -> str:
    return "".join(chars[i] if i < len(chars) else chr(0xE000 + i) for i in ids)

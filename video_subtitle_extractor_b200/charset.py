"""Class-id -> character tables (CTCLabelDecode: ['blank'] + dict lines + [' ']).

The dictionaries ship inside the paddleocr wheel, not in the reference tree.  ``en_dict.txt`` is reconstructible (95
printable ASCII characters, validated by decoding the reference's sample video); other languages need the dictionary
file (``rec_char_dict_path``) — without it the engine still returns class ids and the shim renders them as U+E000-based
private-use code points so that equality/similarity tests between frames (reference backend/main.py:949) keep working.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

EN_DICT = "".join(chr(c) for c in range(0x30, 0x7F)) + "".join(chr(c) for c in range(0x21, 0x30)) + " "


# paddleocr 2.10 resolves `lang` to a dictionary inside its own wheel (paddleocr.py: MODEL_URLS[...]['dict_path'])
_DICT_FILES = {"ch": "ppocr_keys_v1.txt", "en": "en_dict.txt", "chinese_cht": "dict/chinese_cht_dict.txt"}


def find_dict(lang: str) -> Optional[str]:
    """Path of the language's dictionary file if a paddleocr dictionary tree can be found: the directory named by
    VSE_PADDLEOCR_DICT_DIR (= .../paddleocr/ppocr/utils), or a real paddleocr package next to any sys.path entry (the
    compat stand-in of this repo ships none)."""
    import os
    import sys
    rel = _DICT_FILES.get(lang, f"dict/{lang}_dict.txt")
    roots = [os.environ["VSE_PADDLEOCR_DICT_DIR"]] if os.environ.get("VSE_PADDLEOCR_DICT_DIR") else []
    roots += [os.path.join(p, "paddleocr", "ppocr", "utils") for p in sys.path if p]
    for r in roots:
        cand = os.path.join(r, rel)
        if os.path.isfile(cand):
            return cand
    return None


_warned = set()


def characters(lang: str = "en", dict_path: Optional[str] = None, n_classes: Optional[int] = None) -> List[str]:
    if not dict_path and not (lang == "en" or n_classes == 97):
        dict_path = find_dict(lang)
        if dict_path is None and lang not in _warned:
            import warnings
            _warned.add(lang)
            warnings.warn(f"vse_b200: no character dictionary for language '{lang}' (paddleocr ships it inside its wheel; set "
                          "rec_char_dict_path or VSE_PADDLEOCR_DICT_DIR=<paddleocr>/ppocr/utils): recognised text is returned as "
                          "private-use code points U+E000 + class id — comparable between frames, NOT readable", RuntimeWarning,
                          stacklevel=2)
    if dict_path:
        with open(dict_path, "rb") as f:
            lines = [ln.decode("utf-8").strip("\n").strip("\r\n") for ln in f.readlines()]
        return ["blank"] + lines + [" "]
    if lang == "en" or n_classes == 97:
        return ["blank"] + list(EN_DICT[:-1]) + [" "] + [" "]
    n = n_classes or 0x1900
    return ["blank"] + [chr(0xE000 + i) for i in range(1, n)]


def ids_to_text(ids: Sequence[int], chars: Sequence[str]) -> str:
    return "".join(chars[i] if 0 <= i < len(chars) else chr(0xE000 + (i & 0x1FFF)) for i in ids)

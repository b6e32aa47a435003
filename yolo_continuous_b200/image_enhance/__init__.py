from .letter_box import LetterBox, letterbox_batch, letterbox_geometry  # noqa: F401

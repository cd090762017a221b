"""Token sentinels of the reference's prompt protocol (values from
/root/reference/streammind/constants.py:12-13,29-36)."""
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
MMODAL_TOKEN_INDEX = {"IMAGE": -200, "VIDEO": -201, "AUDIO": -202}
MMODAL_INDEX_TOKEN = {v: k for k, v in MMODAL_TOKEN_INDEX.items()}
DEFAULT_MMODAL_TOKEN = {"IMAGE": "<image>", "VIDEO": "<video>", "AUDIO": "<audio>"}

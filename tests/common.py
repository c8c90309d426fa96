import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_vectors():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)["vectors"]


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()

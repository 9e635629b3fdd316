"""Regenerates tests/golden/reference_service_log.json from the reference tree (only where /root/reference exists):
    python tests/golden/make_reference_log_fixture.py [/root/reference]
The one byte-level artefact the reference holds for this path is the response its own client logged at service.log:33 - the first
100 characters of the JSON body, i.e. the first 10 base64 characters of the float32 WAV the engine produced for a 28-character
sentence.  They decode to "RIFF" + the low three bytes of the RIFF size field (+ a zero nibble of the fourth)."""
import base64
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
path = os.path.join(ref, "service.log")
hit = None
for no, line in enumerate(open(path, "rb"), 1):
    m = re.search(rb'"audio_base64": "([A-Za-z0-9+/]+)\.\.\.', line)
    if m:
        hit = (no, m.group(1).decode())
        break
if hit is None:
    raise SystemExit("no audio_base64 prefix in " + path)
line_no, prefix = hit
raw = base64.b64decode(prefix + "A" * (-len(prefix) % 4))[: len(prefix) * 6 // 8]
size_field = int.from_bytes(raw[4:7] + b"\x00", "little")
out = dict(source=f"service.log:{line_no} (first 100 characters of the response body, as the reference's client logged them)",
           audio_base64_prefix=prefix, decoded_hex=raw.hex(), riff_size_field=size_field, file_bytes=size_field + 8,
           note="RIFF size field = file size - 8; only its low three bytes are visible, the next four bits are zero")
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_service_log.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))

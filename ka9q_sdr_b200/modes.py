"""Mode table: the per-mode parameter rows the reference keeps in modes.txt (reference modes.txt:25-39),
parsed there by readmodes (modes.c:32-124) into struct modetab (radio.h:33-48).

In a drop-in deployment the reference's own modes.c/modes.txt stay as they are and hand these values to the
C-ABI (`ka9q_chan_params`); this table is the host-side mirror used by the Python driver, tests and bench.
Parsing rules mirrored from modes.c: low/high are swapped so low <= high (:82-88); attack = -|x| (:90),
recovery = +|x| (:91), hang = |x| (:92); channels default 2, "mono" -> 1 (:96,116-117); "square" implies pll (:110).
"""
from __future__ import annotations

from dataclasses import dataclass

# enum demod_type (reference radio.h:20-24); index into Demodtab (modes.c:25-29)
LINEAR_DEMOD, AM_DEMOD, FM_DEMOD = 0, 1, 2
DEMOD_NAMES = {"LINEAR": LINEAR_DEMOD, "AM": AM_DEMOD, "FM": FM_DEMOD}


@dataclass(frozen=True)
class Mode:
    name: str
    demod_type: int
    low: float
    high: float
    shift: float
    attack: float
    recovery: float
    hang: float
    channels: int = 2
    isb: bool = False
    flat: bool = False
    pll: bool = False
    square: bool = False


def parse_mode_line(line: str) -> Mode | None:
    """Parse one modes.txt row with the rules of modes.c:37-123. Returns None for comments/unknown demods."""
    line = line.split("#", 1)[0].strip()
    tok = line.split()
    if len(tok) < 2:
        return None
    name, demod = tok[0], tok[1].upper()
    dt = None
    for k, v in (("LINEAR", LINEAR_DEMOD), ("AM", AM_DEMOD), ("FM", FM_DEMOD)):
        if demod.startswith(k):  # strncasecmp(demod_name, dtp->name, strlen(dtp->name)) (modes.c:70)
            dt = v
            break
    if dt is None:
        return None
    nums = []
    rest = tok[2:]
    i = 0
    while i < len(rest) and len(nums) < 6:
        try:
            nums.append(float(rest[i]))
        except ValueError:
            break
        i += 1
    while len(nums) < 6:
        nums.append(0.0)
    low, high = nums[0], nums[1]
    if high < low:
        low, high = high, low
    opts = [o.lower() for o in rest[i:i + 8]]
    isb = "isb" in opts or "conj" in opts
    flat = "flat" in opts
    square = "square" in opts
    pll = square or "coherent" in opts or "pll" in opts
    channels = 2
    for o in opts:
        if o == "mono":
            channels = 1
        elif o == "stereo":
            channels = 2
    return Mode(name, dt, low, high, nums[2], -abs(nums[3]), abs(nums[4]), abs(nums[5]), channels, isb, flat, pll,
                square)


# The reference's shipped table (modes.txt:25-39), column for column.
_MODES_TXT = """
FM    FM      -8000  +8000     0    0    0    0
FMF   FM      -8000  +8000     0    0    0    0    flat
AM    AM      -5000  +5000     0  -50  +50  0.0
CAM   LINEAR  -5000  +5000     0  -50  +50  0.0  pll mono
DSB   LINEAR  -5000  +5000     0  -50   +6  1.1  square mono
IQ    LINEAR  -5000  +5000     0  -50   +6  1.1
ISB   LINEAR  -5000  +5000     0  -50   +6  1.1  conj
CISB  LINEAR  -5000  +5000     0  -50   +6  1.1  pll conj
CWU   LINEAR   -200   +200  +700  -50  +20  0.2  mono
CWL   LINEAR   -200   +200  -700  -50  +20  0.2  mono
USB   LINEAR   +100  +3000     0  -50   +6  1.1  mono
LSB   LINEAR  -3000   -100     0  -50   +6  1.1  mono
AME   LINEAR      0  +3000     0  -50  +15  0.0  pll mono
"""

MODES: dict[str, Mode] = {}
for _l in _MODES_TXT.strip().splitlines():
    _m = parse_mode_line(_l)
    if _m is not None:
        MODES[_m.name.upper()] = _m


def get_mode(name: str) -> Mode:
    """Case-insensitive lookup, as set_mode does with strcasecmp (radio.c:327-333)."""
    try:
        return MODES[name.upper()]
    except KeyError:
        raise KeyError(f"unregistered mode {name!r}") from None

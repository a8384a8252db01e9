"""json-backed stand-in for orjson (test tooling only)."""
import json
OPT_INDENT_2 = 1; OPT_SERIALIZE_NUMPY = 2; OPT_NON_STR_KEYS = 4; OPT_SORT_KEYS = 8
def dumps(obj, default=None, option=None):
    return json.dumps(obj, default=default if default else str).encode()
def loads(s):
    return json.loads(s)

"""Extracts the flag table (name, type, default) of the reference entry points' parsers from
/root/reference/utils.py (train_parse utils.py:173-254, sample_parse utils.py:257-327) into
tests/golden/cli_flags.json.  Only the table is extracted (ast walk over the add_argument calls); no code is copied.
Run in the build container, where the read-only reference is mounted:

    python tests/golden/make_cli_fixture.py
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/utils.py'


def flags_of(fn):
    out = {}
    for node in ast.walk(fn):
        if isinstance(node, ast.Call) and getattr(node.func, 'attr', '') == 'add_argument':
            name = node.args[0].value.lstrip('-')
            kw = {k.arg: k.value for k in node.keywords}
            typ = kw['type'].id if 'type' in kw and isinstance(kw['type'], ast.Name) else None
            default = None
            if 'default' in kw:
                try:
                    default = ast.literal_eval(kw['default'])
                except Exception:
                    default = '<expr>'          # computed default (save_dir, tool directories)
            out[name] = {'type': typ, 'default': default}
    return out


def main():
    # the file is Python 2 (print statements elsewhere): parse the two parser functions on their own
    src = open(REF).read().split('\n')
    starts = [i for i, l in enumerate(src) if l.startswith('def ')] + [len(src)]
    table = {}
    for a, b in zip(starts[:-1], starts[1:]):
        name = src[a][4:src[a].index('(')]
        if name in ('train_parse', 'sample_parse'):
            table[name] = flags_of(ast.parse('\n'.join(src[a:b])))
    json.dump(table, open(os.path.join(HERE, 'cli_flags.json'), 'w'), indent=1, sort_keys=True)
    print({k: len(v) for k, v in table.items()})


if __name__ == '__main__':
    main()

"""Console helpers with the reference's names, prefixes and verbosity gating
(reference: bin/sofa_print.py).  Output strings are part of the observable surface
(`test/test.py:72` greps stdout), so prefixes and ANSI codes are kept."""

_ESC = '\033['
_END = _ESC + '0m'
_STYLE = {'title': _ESC + '48;5;26m', 'error': _ESC + '1;31m', 'warning': _ESC + '93m',
          'info': _ESC + '92m', 'hint': _ESC + '38;5;45m', 'progress': _ESC + '38;5;104m'}


def _tagged(style, tag, content):
    print(_STYLE[style] + tag + content + _END)


def print_title(content):
    print('\n')
    print(_STYLE['title'] + content + _END)


def print_error(content):
    _tagged('error', "[ERROR] ", content)


def print_warning(cfg, content):
    if cfg.verbose:
        _tagged('warning', "[WARNING] ", content)


def print_info(cfg, content):
    if cfg.verbose:
        _tagged('info', "[INFO] ", content)


def print_hint(content):
    _tagged('hint', "[HINT] ", content)


def print_progress(cfg, content):
    if cfg.verbose:
        _tagged('progress', "[PROGRESS] ", content)


def print_main_progress(content):
    _tagged('progress', "[PROGRESS] ", content)

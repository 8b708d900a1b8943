"""Build `tests/emu/_build/libsfb_emu.so`: the N3 kernel sources compiled for the CPU SIMT emulator (tests/emu/common.cuh).

TEST INFRASTRUCTURE ONLY.  The sources are taken from synchformer_b200/csrc/ as they are; the only textual rewrites are the two that g++
cannot express through macros:
  * `kernel<<<grid, block, smem, stream>>>(args);`  ->  `emu::launch(dim3(grid), dim3(block), smem, [&]() { kernel(args); });`
  * `extern __shared__ ... smem_raw[];`             ->  `unsigned char *smem_raw = emu::dyn_smem();`
`#include "common.cuh"` resolves to the shim next to the generated copy; `philox.cuh` is the real one (via -I csrc).
"""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, 'synchformer_b200', 'csrc')
OUT = os.path.join(HERE, '_build')
LIB = os.path.join(OUT, 'libsfb_emu.so')
SOURCES = ['train.cu', 'attention_train.cu', 'attention_bwd.cu', 'optim.cu', 'contrastive.cu',
           # kernels that were verified on the B200 in round 1 and contain no inline PTX: running them here validates the emulator itself
           'layernorm.cu', 'embed.cu', 'attention.cu']
CUDA_INC = os.environ.get('CUDA_INC', '/usr/local/cuda/include')


def _match(text: str, start: int, open_ch: str, close_ch: str) -> int:
    """index just past the bracket that closes the one at text[start]"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError('unbalanced')


def _split_top(s: str):
    parts, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{<':
            depth += 1
        elif ch in ')]}>':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur.strip())
            cur = ''
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text: str) -> str:
    out, pos = '', 0
    while True:
        k = text.find('<<<', pos)
        if k < 0:
            return out + text[pos:]
        name_start = k
        if text[name_start - 1] == '>':                       # template arguments (may contain spaces): back to the matching '<'
            depth = 0
            while True:
                name_start -= 1
                depth += {'>': 1, '<': -1}.get(text[name_start], 0)
                if depth == 0:
                    break
        while name_start > 0 and re.match(r'[A-Za-z0-9_:]', text[name_start - 1]):
            name_start -= 1
        cfg_end = text.index('>>>', k)
        cfg = _split_top(text[k + 3:cfg_end])
        assert 2 <= len(cfg) <= 4, cfg
        args_start = cfg_end + 3
        assert text[args_start] == '(', text[args_start:args_start + 20]
        args_end = _match(text, args_start, '(', ')')
        kernel, args = text[name_start:k], text[args_start:args_end]
        smem = cfg[2] if len(cfg) > 2 else '0'
        out += text[pos:name_start] + f'emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), {smem}, [&]() {{ {kernel}{args}; }})'
        pos = args_end
    return out


PTX = [  # (substring of the PTX text, emulator function, uses outputs)
    ('cp.async.cg.shared.global', 'emu::ptx_cp_async16'), ('cp.async.commit_group', 'emu::ptx_nop'), ('cp.async.wait_group', 'emu::ptx_nop'),
    ('ldmatrix.sync.aligned.m8n8.x4.trans', 'emu::ptx_ldmatrix_x4_trans'), ('ldmatrix.sync.aligned.m8n8.x4', 'emu::ptx_ldmatrix_x4'),
    ('ldmatrix.sync.aligned.m8n8.x2', 'emu::ptx_ldmatrix_x2'), ('mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32', 'emu::ptx_mma_16816'),
    ('ex2.approx.ftz.f32', 'emu::ptx_ex2'),
]


def rewrite_asm(text: str) -> str:
    """`asm [volatile]("ptx" : outs : ins : clobbers);` -> `emu::ptx_xxx(out expressions..., in expressions...);` for the instructions in PTX."""
    out, pos = '', 0
    for m in re.finditer(r'\basm\s*(?:volatile\s*)?\(', text):
        if m.start() < pos:
            continue
        end = _match(text, m.end() - 1, '(', ')')
        body = text[m.end():end - 1]
        strings = re.findall(r'"((?:[^"\\]|\\.)*)"', body.split(':')[0] if ':' in body else body)
        ptx = ''.join(strings)
        sections, depth, cur, in_str = [], 0, '', False
        for ch in body:
            if ch == '"':
                in_str = not in_str
            if not in_str:
                depth += {'(': 1, ')': -1}.get(ch, 0)
                if ch == ':' and depth == 0:
                    sections.append(cur)
                    cur = ''
                    continue
            cur += ch
        sections.append(cur)
        operands = []
        for sec in sections[1:3]:
            for op in _split_top(sec):
                mm = re.match(r'\s*"[^"]*"\s*\((.*)\)\s*$', op, re.S)
                if mm:
                    operands.append(mm.group(1).strip())
        fn = next((f for key, f in PTX if key in ptx), None)
        if fn is None:
            raise ValueError(f'no emulation for PTX: {ptx[:80]}')
        if fn == 'emu::ptx_nop':
            operands = []
        out += text[pos:m.start()] + f'{fn}({", ".join(operands)})'
        pos = end
    return out + text[pos:]


def transform(text: str) -> str:
    if 'asm' in text:
        text = rewrite_asm(text)
    text = re.sub(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?((?:unsigned char)|\w+)\s+(\w+)\[\];',
                  r'\1 *\2 = reinterpret_cast<\1 *>(emu::dyn_smem());', text)
    return rewrite_launches(text)


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for f in SOURCES + ['philox.cuh', 'attention.cuh']:
        h.update(open(os.path.join(CSRC, f), 'rb').read())
    for f in ('common.cuh', 'build_emu.py'):
        h.update(open(os.path.join(HERE, f), 'rb').read())
    stamp = os.path.join(OUT, 'stamp')
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    gen = []
    for f in SOURCES:
        dst = os.path.join(OUT, f.replace('.cu', '_emu.cpp'))
        with open(dst, 'w') as fh:
            fh.write(transform(open(os.path.join(CSRC, f)).read()))
        gen.append(dst)
    with open(os.path.join(OUT, 'common.cuh'), 'w') as fh:          # quote-includes look next to the including file first
        fh.write('#include "../common.cuh"\n')
    with open(os.path.join(OUT, 'attention.cuh'), 'w') as fh:       # same header without <cuda_runtime.h> (the shim provides the types)
        fh.write(open(os.path.join(CSRC, 'attention.cuh')).read().replace('#include <cuda_runtime.h>', ''))
    extra = os.path.join(OUT, 'emu_exports.cpp')
    with open(extra, 'w') as fh:
        fh.write('#define EMU_MAIN_TU 1\n#include "common.cuh"\nextern "C" const char *sfb_last_error(void) { return sfb::err_buf(); }\n'
                 'extern "C" long emu_launch_count(void) { return emu::S().launches; }\n'
                 '#include "attention.cuh"\n'        # the tcgen05 space-attention kernel is not emulated: report it as unsupported
                 'namespace sfb { namespace attn { bool tc_supported(const Desc &) { return false; } '
                 'int launch_tc(const Desc &, cudaStream_t) { return SFB_E_UNSUPPORTED; } } }\n'
                 'extern "C" void emu_set_schedule(int mode, unsigned long long seed) { emu::schedule() = mode; emu::rng() = seed * 2 + 1; }\n')
    cmd = ['g++', '-O2', '-g', '-std=c++17', '-shared', '-fPIC', '-w', '-fno-strict-aliasing',   # CUDA code type-puns freely; nvcc does not assume strict aliasing
           '-I', OUT, '-I', CSRC, '-I', CUDA_INC, '-o', LIB] + gen + [extra]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('emulator build failed:\n' + r.stdout + r.stderr)
    with open(stamp, 'w') as fh:
        fh.write(h.hexdigest())
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv))

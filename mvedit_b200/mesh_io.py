"""On-disk formats of the mesh stage's result: ``Mesh.write(path)`` for ``.obj`` (+ ``.mtl`` + albedo PNG), ``.ply``
(geometry only) and ``.glb`` (geometry + UVs + normals + embedded albedo PNG), what the runner hands back to the web UI
(``lib/models/decoders/mesh_renderer/mesh_utils.py:461-692``).  Host-side I/O, no GPU work: tensors are copied to the CPU once.

The reference goes through trimesh (``.ply``), pygltflib (``.glb``) and cv2 (PNG); none of the first two is installed here, so the
containers are written by hand: binary little-endian PLY, and glTF 2.0 binary (12-byte header, JSON chunk, BIN chunk, 4-byte aligned)
with the same accessor layout the reference builds (indices, POSITION, TEXCOORD_0, NORMAL, one embedded ``image/png``; linear /
mip-mapped sampler, REPEAT wrap, metallic 0 / roughness 1, double sided).  Like the reference, ``.obj`` flips v to ``1 - v``.
"""
import io
import json
import os
import struct

import numpy as np
import torch


def _np(t, dtype=None):
    a = t.detach().cpu().numpy()
    return a if dtype is None else a.astype(dtype)


def _png_bytes(albedo_hwc_float):
    from PIL import Image
    img = (np.clip(albedo_hwc_float[..., :3], 0, 1) * 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img, 'RGB').save(buf, format='PNG')
    return buf.getvalue()


def align_to_vt(mesh):
    """One vertex per UV vertex (``Mesh.align_v_to_vt`` / ``align_vn_to_vt``, mesh_utils.py:416-438): glTF has a single index buffer, so
    positions / normals are re-indexed by the texture topology (a position shared by several UV vertices is duplicated)."""
    ft, f = mesh.ft.reshape(-1).long(), mesh.f.reshape(-1).long()
    vmap = torch.zeros(mesh.vt.shape[0], dtype=torch.long, device=mesh.v.device)
    vmap[ft] = f
    nmap = torch.zeros_like(vmap)
    nmap[ft] = mesh.fn.reshape(-1).long()
    return mesh.v[vmap], mesh.vn[nmap], mesh.vt, mesh.ft


def write_obj(mesh, path):
    mtl_path, albedo_path = path.replace('.obj', '.mtl'), path.replace('.obj', '_albedo.png')
    v, f = _np(mesh.v), _np(mesh.f)
    vt, vn = (None if mesh.vt is None else _np(mesh.vt)), (None if mesh.vn is None else _np(mesh.vn))
    ft, fn = (None if mesh.ft is None else _np(mesh.ft)), (None if mesh.fn is None else _np(mesh.fn))
    textured = (not mesh.textureless) and mesh.albedo is not None
    lines = ['mtllib %s' % os.path.basename(mtl_path)]
    lines += ['v %.6f %.6f %.6f' % tuple(p) for p in v]
    if vt is not None:
        lines += ['vt %.4f %.4f' % (p[0], 1 - p[1]) for p in vt]
    if vn is not None:
        lines += ['vn %.4f %.4f %.4f' % tuple(p) for p in vn]
    lines.append('usemtl defaultMat')
    corner = lambda i, k: '%d/%s/%s' % (f[i, k] + 1, '' if ft is None else ft[i, k] + 1, '' if fn is None else fn[i, k] + 1)
    lines += ['f %s %s %s' % (corner(i, 0), corner(i, 1), corner(i, 2)) for i in range(len(f))]
    with open(path, 'w') as fp:
        fp.write('\n'.join(lines) + '\n')
    with open(mtl_path, 'w') as fp:
        fp.write('newmtl defaultMat\nKa 1 1 1\nKd 1 1 1\nKs 0 0 0\nTr 1\nillum 1\nNs 0\n' + ('map_Kd %s\n' % os.path.basename(albedo_path) if textured else ''))
    if textured:
        with open(albedo_path, 'wb') as fp:
            fp.write(_png_bytes(_np(mesh.albedo)))


def write_ply(mesh, path):
    v, f = _np(mesh.v, np.float32), _np(mesh.f, np.int32)
    header = ('ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n'
              'element face %d\nproperty list uchar int vertex_indices\nend_header\n' % (len(v), len(f)))
    rec = np.empty(len(f), dtype=[('n', 'u1'), ('i', '<i4', 3)])
    rec['n'], rec['i'] = 3, f
    with open(path, 'wb') as fp:
        fp.write(header.encode('ascii'))
        fp.write(v.astype('<f4').tobytes())
        fp.write(rec.tobytes())


def write_glb(mesh, path):
    assert mesh.vn is not None, 'write_glb needs vertex normals (auto_normal)'
    if mesh.vt is None:
        v, vn, vt, f = mesh.v, mesh.vn, mesh.v.new_zeros((mesh.v.size(0), 2)), mesh.f
    else:
        v, vn, vt, f = align_to_vt(mesh)
    blobs = [_np(f, np.uint32).reshape(-1).tobytes(), _np(v, np.float32).tobytes(), _np(vt, np.float32).tobytes(), _np(vn, np.float32).tobytes(),
             _png_bytes(_np(mesh.albedo) if mesh.albedo is not None else np.full((1024, 1024, 3), 0.5, np.float32))]
    views, off = [], 0
    for k, b in enumerate(blobs):
        view = dict(buffer=0, byteOffset=off, byteLength=len(b))
        if k == 0:
            view['target'] = 34963                      # ELEMENT_ARRAY_BUFFER
        elif k < 4:
            view['target'] = 34962                      # ARRAY_BUFFER
            view['byteStride'] = (12, 8, 12)[k - 1]
        views.append(view)
        off += len(b) + (-len(b)) % 4
    vmin, vmax = _np(v, np.float32).min(0).tolist(), _np(v, np.float32).max(0).tolist()
    n_v, n_i = int(v.shape[0]), int(f.numel())
    gltf = dict(
        asset=dict(version='2.0', generator='mvedit_b200'), scene=0, scenes=[dict(nodes=[0])], nodes=[dict(mesh=0)],
        meshes=[dict(primitives=[dict(attributes=dict(POSITION=1, TEXCOORD_0=2, NORMAL=3), indices=0, material=0)])],
        materials=[dict(pbrMetallicRoughness=dict(baseColorTexture=dict(index=0, texCoord=0), metallicFactor=0.0, roughnessFactor=1.0),
                        alphaCutoff=0, doubleSided=True)],
        textures=[dict(sampler=0, source=0)], samplers=[dict(magFilter=9729, minFilter=9987, wrapS=10497, wrapT=10497)],
        images=[dict(bufferView=4, mimeType='image/png')], buffers=[dict(byteLength=off)], bufferViews=views,
        accessors=[dict(bufferView=0, componentType=5125, count=n_i, type='SCALAR', max=[int(_np(f).max())], min=[int(_np(f).min())]),
                   dict(bufferView=1, componentType=5126, count=n_v, type='VEC3', max=vmax, min=vmin),
                   dict(bufferView=2, componentType=5126, count=n_v, type='VEC2', max=_np(vt, np.float32).max(0).tolist(), min=_np(vt, np.float32).min(0).tolist()),
                   dict(bufferView=3, componentType=5126, count=n_v, type='VEC3', max=_np(vn, np.float32).max(0).tolist(), min=_np(vn, np.float32).min(0).tolist())])
    js = json.dumps(gltf, separators=(',', ':')).encode('utf-8')
    js += b' ' * ((-len(js)) % 4)
    bin_chunk = b''.join(b + b'\x00' * ((-len(b)) % 4) for b in blobs)
    total = 12 + 8 + len(js) + 8 + len(bin_chunk)
    with open(path, 'wb') as fp:
        fp.write(struct.pack('<4sII', b'glTF', 2, total))
        fp.write(struct.pack('<I4s', len(js), b'JSON') + js)
        fp.write(struct.pack('<I4s', len(bin_chunk), b'BIN\x00') + bin_chunk)


def write(mesh, path, flip_yz=False):
    """``Mesh.write`` (mesh_utils.py:461-477): ``flip_yz`` swaps to a y-up frame (y <- z, z <- -y)."""
    if flip_yz:
        mesh = mesh.copy()
        flip = lambda t: torch.stack([t[..., 0], t[..., 2], -t[..., 1]], dim=-1)
        mesh.v, mesh.vn = flip(mesh.v), (None if mesh.vn is None else flip(mesh.vn))
    if path.endswith('.ply'):
        write_ply(mesh, path)
    elif path.endswith('.obj'):
        write_obj(mesh, path)
    elif path.endswith('.glb') or path.endswith('.gltf'):
        write_glb(mesh, path)
    else:
        raise NotImplementedError('format %s not supported!' % path)


# ---- loading (mesh_utils.py:80-260): .obj with face uvs / normals and a map_Kd texture, and the binary .ply written above ------------------

def load_obj(path, device=None):
    from .mesh_renderer import Mesh
    v, vt, vn, f, ft, fn, mtl = [], [], [], [], [], [], None
    with open(path) as fp:
        for line in fp:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'v':
                v.append([float(x) for x in tok[1:4]])
            elif tok[0] == 'vt':
                vt.append([float(tok[1]), 1.0 - float(tok[2])])                 # stored as 1 - v (write_obj)
            elif tok[0] == 'vn':
                vn.append([float(x) for x in tok[1:4]])
            elif tok[0] == 'mtllib':
                mtl = tok[1]
            elif tok[0] == 'f':
                corners = [(c.split('/') + ['', ''])[:3] for c in tok[1:]]
                idx = [[int(c[k]) - 1 if c[k] else -1 for c in corners] for k in range(3)]
                for j in range(1, len(corners) - 1):                            # fan triangulation of polygons
                    f.append([idx[0][0], idx[0][j], idx[0][j + 1]])
                    ft.append([idx[1][0], idx[1][j], idx[1][j + 1]])
                    fn.append([idx[2][0], idx[2][j], idx[2][j + 1]])
    t = lambda a, dt: torch.tensor(a, dtype=dt, device=device)
    mesh = Mesh(v=t(v, torch.float32), f=t(f, torch.int32), device=device)
    if vt and min(min(r) for r in ft) >= 0:
        mesh.vt, mesh.ft = t(vt, torch.float32), t(ft, torch.int32)
    if vn and min(min(r) for r in fn) >= 0:
        mesh.vn, mesh.fn = t(vn, torch.float32), t(fn, torch.int32)
    if mtl is not None:
        mtl_path = os.path.join(os.path.dirname(path), mtl)
        if os.path.exists(mtl_path):
            for line in open(mtl_path):
                tok = line.split()
                if tok and tok[0] == 'map_Kd':
                    from PIL import Image
                    img = np.asarray(Image.open(os.path.join(os.path.dirname(path), tok[1])).convert('RGB')).astype(np.float32) / 255
                    mesh.albedo = torch.cat([torch.from_numpy(img), torch.ones(img.shape[0], img.shape[1], 1)], dim=-1).to(device)
    mesh.textureless = mesh.albedo is None
    return mesh


def load_ply(path, device=None):
    from .mesh_renderer import Mesh
    raw = open(path, 'rb').read()
    head, body = raw.split(b'end_header\n', 1)
    if b'binary_little_endian' not in head:
        raise NotImplementedError('load_ply: only the binary little-endian layout written by write_ply is read')
    nv = int([l for l in head.split(b'\n') if l.startswith(b'element vertex')][0].split()[-1])
    nf = int([l for l in head.split(b'\n') if l.startswith(b'element face')][0].split()[-1])
    v = np.frombuffer(body[:nv * 12], '<f4').reshape(nv, 3)
    faces = np.frombuffer(body[nv * 12:nv * 12 + nf * 13], dtype=[('n', 'u1'), ('i', '<i4', 3)])
    return Mesh(v=torch.from_numpy(v.copy()).to(device), f=torch.from_numpy(faces['i'].copy()).to(device), device=device, textureless=True)


_GLTF_DTYPES = {5120: 'i1', 5121: 'u1', 5122: '<i2', 5123: '<u2', 5125: '<u4', 5126: '<f4'}
_GLTF_NCOMP = {'SCALAR': 1, 'VEC2': 2, 'VEC3': 3, 'VEC4': 4, 'MAT4': 16}


def load_glb(path, device=None):
    """Binary glTF 2.0 with ONE mesh of one triangle primitive -- what ``Mesh.load_trimesh`` accepts (mesh_utils.py:262-345; the runner's
    3D-to-3D inputs are the ``.glb`` files the pipelines write): positions, indices, optional normals / TEXCOORD_0 / COLOR_0 and the
    base-colour texture.  As the reference (which reads ``geometry[key]`` of the trimesh scene) the node transforms are not applied, and
    ``vt`` is the file's TEXCOORD_0 as stored (trimesh flips v on load, the reference flips it back, :301-302)."""
    from .mesh_renderer import Mesh
    raw = open(path, 'rb').read()
    magic, version, total = struct.unpack_from('<4sII', raw, 0)
    if magic != b'glTF' or version != 2:
        raise ValueError('%s is not a binary glTF 2.0 file' % path)
    off, js, bin_chunk = 12, None, b''
    while off + 8 <= min(total, len(raw)):
        n, kind = struct.unpack_from('<I4s', raw, off)
        body = raw[off + 8:off + 8 + n]
        if kind == b'JSON':
            js = json.loads(body.decode('utf-8'))
        elif kind == b'BIN\x00':
            bin_chunk = body
        off += 8 + n
    meshes = js.get('meshes', [])
    if len(meshes) != 1 or len(meshes[0]['primitives']) != 1:
        raise NotImplementedError('%s contains more than one mesh / primitive, not supported!' % path)
    prim = meshes[0]['primitives'][0]
    if prim.get('mode', 4) != 4:
        raise NotImplementedError('load_glb: only triangle lists (mode 4) are read')

    def view_bytes(i):
        bv = js['bufferViews'][i]
        if bv.get('buffer', 0) != 0 or 'uri' in js['buffers'][0]:
            raise NotImplementedError('load_glb: external buffers are not read')
        return bin_chunk[bv.get('byteOffset', 0):bv.get('byteOffset', 0) + bv['byteLength']], bv.get('byteStride')

    def accessor(i):
        acc = js['accessors'][i]
        data, stride = view_bytes(acc['bufferView'])
        dt, nc = np.dtype(_GLTF_DTYPES[acc['componentType']]), _GLTF_NCOMP[acc['type']]
        start, count, item = acc.get('byteOffset', 0), acc['count'], dt.itemsize * nc
        if stride in (None, 0, item):
            arr = np.frombuffer(data, dt, count * nc, start).reshape(count, nc)
        else:
            arr = np.stack([np.frombuffer(data, dt, nc, start + k * stride) for k in range(count)])
        if acc.get('normalized', False) and dt.kind in 'iu':
            arr = arr.astype(np.float32) / np.iinfo(dt).max
        return arr
    attr = prim['attributes']
    v = accessor(attr['POSITION']).astype(np.float32)
    f = accessor(prim['indices']).reshape(-1, 3).astype(np.int32) if 'indices' in prim else np.arange(len(v), dtype=np.int32).reshape(-1, 3)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
    mesh = Mesh(v=t(v, torch.float32), f=t(f, torch.int32), device=device)
    if 'NORMAL' in attr:
        mesh.vn, mesh.fn = t(accessor(attr['NORMAL']).astype(np.float32), torch.float32), mesh.f
    if 'TEXCOORD_0' in attr:
        mesh.vt, mesh.ft = t(accessor(attr['TEXCOORD_0']).astype(np.float32), torch.float32), mesh.f
    if 'COLOR_0' in attr:
        col = accessor(attr['COLOR_0'])
        col = col.astype(np.float32) / (np.iinfo(col.dtype).max if col.dtype.kind in 'iu' else 1.0)
        mesh.vc = t(np.concatenate([col, np.ones((len(col), 1), np.float32)], axis=1) if col.shape[1] == 3 else col, torch.float32)
    tex = js.get('materials', [{}])[prim.get('material', 0)].get('pbrMetallicRoughness', {}).get('baseColorTexture') if js.get('materials') else None
    if tex is not None:
        import io
        from PIL import Image
        img = js['images'][js['textures'][tex['index']]['source']]
        if 'bufferView' in img:
            blob = view_bytes(img['bufferView'])[0]
        elif img.get('uri', '').startswith('data:'):
            import base64
            blob = base64.b64decode(img['uri'].split(',', 1)[1])
        else:
            blob = open(os.path.join(os.path.dirname(path), img['uri']), 'rb').read()
        mesh.albedo = t(np.asarray(Image.open(io.BytesIO(blob))).astype(np.float32) / 255, torch.float32)
    mesh.textureless = mesh.albedo is None
    return mesh


def load(path, resize=False, auto_uv=True, flip_yz=False, force_auto_normal=False, auto_normal_seamless=False, device=None, mesh=None):
    """``Mesh.load`` (mesh_utils.py:80-113): read (or take ``mesh``, the reference's ``path=None`` + constructor kwargs), fix normals / UVs,
    optional y-up -> z-up flip (the inverse of ``write``'s)."""
    if mesh is not None:
        pass
    elif path.endswith('.obj'):
        mesh = load_obj(path, device)
    elif path.endswith('.ply'):
        mesh = load_ply(path, device)
    elif path.endswith('.glb'):
        mesh = load_glb(path, device)
    else:
        raise NotImplementedError('Mesh.load: %s -- .obj, .glb and the binary .ply of Mesh.write are read' % path)
    if resize:
        vmin, vmax = mesh.v.min(dim=0).values, mesh.v.max(dim=0).values
        mesh.ori_center, mesh.ori_scale = (vmax + vmin) / 2, 1.2 / float((vmax - vmin).max())
        mesh.v = (mesh.v - mesh.ori_center) * mesh.ori_scale
    if mesh.vn is None or force_auto_normal:
        mesh.auto_normal(seamless=auto_normal_seamless)
    if mesh.vt is None and auto_uv:
        mesh.auto_uv()
    if flip_yz:
        flip = lambda t: torch.stack([t[..., 0], -t[..., 2], t[..., 1]], dim=-1)
        mesh.v, mesh.vn = flip(mesh.v), flip(mesh.vn)
    return mesh

"""Mesh.write: .obj / .ply / .glb containers parsed back on the CPU."""
import io
import json
import struct

import numpy as np
import torch

from tests import synth_mesh
from mvedit_b200.mesh_renderer import Mesh


def _mesh(textured=True):
    v, f = synth_mesh.icosphere(1)
    m = Mesh(v=torch.from_numpy(v).float() * 0.5, f=torch.from_numpy(f).int())
    m.auto_normal()
    if textured:
        m.auto_uv()
        m.albedo = torch.rand(16, 16, 4, generator=torch.Generator().manual_seed(0))
    return m


def test_obj_round_trip(tmp_path):
    m = _mesh()
    p = str(tmp_path / 'a.obj')
    m.write(p)
    lines = open(p).read().splitlines()
    vs = np.array([[float(x) for x in l.split()[1:]] for l in lines if l.startswith('v ')])
    vts = np.array([[float(x) for x in l.split()[1:]] for l in lines if l.startswith('vt ')])
    fs = [l.split()[1:] for l in lines if l.startswith('f ')]
    np.testing.assert_allclose(vs, m.v.numpy(), atol=1e-6)
    np.testing.assert_allclose(vts, np.stack([m.vt[:, 0].numpy(), 1 - m.vt[:, 1].numpy()], -1), atol=1e-4)      # v flipped, as the reference
    assert len(fs) == m.f.shape[0] and fs[0][0] == '%d/%d/%d' % (m.f[0, 0] + 1, m.ft[0, 0] + 1, m.fn[0, 0] + 1)
    assert 'map_Kd a_albedo.png' in open(str(tmp_path / 'a.mtl')).read()
    from PIL import Image
    img = np.asarray(Image.open(str(tmp_path / 'a_albedo.png')))
    assert img.shape == (16, 16, 3) and np.abs(img.astype(np.float32) / 255 - m.albedo[..., :3].numpy()).max() < 1 / 255 + 1e-6
    # untextured: no vt, no texture reference
    m2 = _mesh(False)
    m2.write(str(tmp_path / 'b.obj'))
    assert 'map_Kd' not in open(str(tmp_path / 'b.mtl')).read() and '//' in open(str(tmp_path / 'b.obj')).read()


def test_ply_and_yz_flip(tmp_path):
    m = _mesh(False)
    p = str(tmp_path / 'a.ply')
    m.write(p, flip_yz=True)
    raw = open(p, 'rb').read()
    head, body = raw.split(b'end_header\n', 1)
    assert b'element vertex %d' % m.v.shape[0] in head and b'element face %d' % m.f.shape[0] in head
    v = np.frombuffer(body[:m.v.shape[0] * 12], '<f4').reshape(-1, 3)
    np.testing.assert_allclose(v, np.stack([m.v[:, 0], m.v[:, 2], -m.v[:, 1]], -1), atol=1e-7)
    faces = np.frombuffer(body[m.v.shape[0] * 12:], dtype=[('n', 'u1'), ('i', '<i4', 3)])
    assert (faces['n'] == 3).all() and (faces['i'] == m.f.numpy()).all()
    assert torch.equal(m.v, _mesh(False).v)                   # the flip works on a copy


def test_glb_container(tmp_path):
    m = _mesh()
    p = str(tmp_path / 'a.glb')
    m.write(p)
    raw = open(p, 'rb').read()
    magic, version, total = struct.unpack('<4sII', raw[:12])
    assert magic == b'glTF' and version == 2 and total == len(raw)
    jlen, jtype = struct.unpack('<I4s', raw[12:20])
    gltf = json.loads(raw[20:20 + jlen])
    blen, btype = struct.unpack('<I4s', raw[20 + jlen:28 + jlen])
    assert jtype == b'JSON' and btype == b'BIN\x00' and jlen % 4 == 0 and blen % 4 == 0 and 28 + jlen + blen == len(raw)
    bin_ = raw[28 + jlen:]
    acc, views = gltf['accessors'], gltf['bufferViews']
    n_uv = m.vt.shape[0]
    assert acc[0]['count'] == m.f.numel() and acc[1]['count'] == acc[2]['count'] == acc[3]['count'] == n_uv
    idx = np.frombuffer(bin_[views[0]['byteOffset']:views[0]['byteOffset'] + views[0]['byteLength']], '<u4').reshape(-1, 3)
    pos = np.frombuffer(bin_[views[1]['byteOffset']:views[1]['byteOffset'] + views[1]['byteLength']], '<f4').reshape(-1, 3)
    np.testing.assert_allclose(pos[idx], m.v.numpy()[m.f.numpy()], atol=1e-7)          # re-indexed by the UV topology, same triangles
    from PIL import Image
    png = bin_[views[4]['byteOffset']:views[4]['byteOffset'] + views[4]['byteLength']]
    assert np.asarray(Image.open(io.BytesIO(png))).shape == (16, 16, 3)
    assert gltf['images'][0]['mimeType'] == 'image/png' and gltf['materials'][0]['doubleSided'] is True


def test_obj_and_ply_load_back(tmp_path):
    m = _mesh()
    p = str(tmp_path / 'a.obj')
    m.write(p, flip_yz=True)
    r = Mesh.load(p, flip_yz=True)                                     # the load-side flip undoes the write-side one
    np.testing.assert_allclose(r.v.numpy(), m.v.numpy(), atol=2e-6)
    np.testing.assert_allclose(r.vt.numpy(), m.vt.numpy(), atol=1e-4)
    np.testing.assert_allclose(r.vn.numpy(), m.vn.numpy(), atol=1e-4)
    assert torch.equal(r.f, m.f) and torch.equal(r.ft, m.ft) and torch.equal(r.fn, m.fn) and r.textureless is False
    assert r.albedo.shape == (16, 16, 4) and (r.albedo[..., :3] - m.albedo[..., :3]).abs().max() < 1 / 255 + 1e-6
    q = str(tmp_path / 'b.ply')
    m.write(q)
    g = Mesh.load(q)
    assert torch.equal(g.f, m.f) and torch.allclose(g.v, m.v) and g.vn is not None and g.vt is not None      # normals / atlas rebuilt
    (tmp_path / 'quad.obj').write_text('v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n')
    quad = Mesh.load(str(tmp_path / 'quad.obj'), auto_uv=False)
    assert quad.f.tolist() == [[0, 1, 2], [0, 2, 3]] and quad.vt is None


def test_glb_load_back(tmp_path):
    """``Mesh.load`` of a binary glTF (the reference goes through trimesh, mesh_utils.py:262-345): the file the writer produced comes back
    with the same triangles, uvs, normals and texture; a hand-built file with strided / normalised accessors, vertex colours and no
    indices is decoded too."""
    m = _mesh()
    p = str(tmp_path / 'a.glb')
    m.write(p, flip_yz=True)
    r = Mesh.load(p, flip_yz=True)
    np.testing.assert_allclose(r.v.numpy()[r.f.numpy()], m.v.numpy()[m.f.numpy()], atol=1e-6)              # same triangles (vertices re-indexed by the uvs)
    np.testing.assert_allclose(r.vt.numpy()[r.ft.numpy()], m.vt.numpy()[m.ft.numpy()], atol=1e-6)
    np.testing.assert_allclose(r.vn.numpy()[r.fn.numpy()], m.vn.numpy()[m.fn.numpy()], atol=1e-5)
    assert r.textureless is False and r.albedo.shape == (16, 16, 3) and (r.albedo - m.albedo[..., :3]).abs().max() < 1 / 255 + 1e-6
    # hand-built: interleaved POSITION (stride 16), normalised u16 TEXCOORD_0, u8 COLOR_0, no indices, no material
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    inter = np.zeros((6, 4), np.float32)
    inter[:, :3] = pos
    uv = (np.array([[0, 0], [1, 0], [0, 1], [1, 0], [1, 1], [0, 1]], np.float32) * 65535).astype('<u2')
    col = np.array([[255, 0, 0, 255]] * 3 + [[0, 255, 0, 128]] * 3, np.uint8)
    blobs = [inter.tobytes(), uv.tobytes(), col.tobytes()]
    views, off = [], 0
    for k, b in enumerate(blobs):
        views.append(dict(buffer=0, byteOffset=off, byteLength=len(b), **({'byteStride': 16} if k == 0 else {})))
        off += len(b) + (-len(b)) % 4
    gltf = dict(asset=dict(version='2.0'), meshes=[dict(primitives=[dict(attributes=dict(POSITION=0, TEXCOORD_0=1, COLOR_0=2))])],
                buffers=[dict(byteLength=off)], bufferViews=views,
                accessors=[dict(bufferView=0, componentType=5126, count=6, type='VEC3'),
                           dict(bufferView=1, componentType=5123, count=6, type='VEC2', normalized=True),
                           dict(bufferView=2, componentType=5121, count=6, type='VEC4', normalized=True)])
    js = json.dumps(gltf).encode()
    js += b' ' * ((-len(js)) % 4)
    bin_ = b''.join(b + b'\x00' * ((-len(b)) % 4) for b in blobs)
    q = str(tmp_path / 'hand.glb')
    with open(q, 'wb') as fp:
        fp.write(struct.pack('<4sII', b'glTF', 2, 28 + len(js) + len(bin_)) + struct.pack('<I4s', len(js), b'JSON') + js
                 + struct.pack('<I4s', len(bin_), b'BIN\x00') + bin_)
    h = Mesh.load(q, auto_uv=False)
    assert h.f.tolist() == [[0, 1, 2], [3, 4, 5]] and h.textureless is True and h.vn is not None
    np.testing.assert_allclose(h.v.numpy(), pos)
    np.testing.assert_allclose(h.vt.numpy(), uv.astype(np.float32) / 65535)
    np.testing.assert_allclose(h.vc.numpy(), col.astype(np.float32) / 255)

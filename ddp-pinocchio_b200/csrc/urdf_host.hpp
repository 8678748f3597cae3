// urdf_host.hpp -- URDF text -> model description (host side, pure C++, no CUDA, no XML library).
//
// Replaces pinocchio::urdf::buildModel as the reference uses it (/root/reference/include/ddp/pinocchio_model.ipp:98-113:
// model_t{urdf_path} -> buildModel(path, model), fixed base) for the models this library supports: serial chains of
// revolute / continuous / prismatic joints (one degree of freedom each, nq = nv).  A continuous joint is taken as a revolute
// joint without limits on the plain angle; Pinocchio's parser gives it the (cos, sin) parametrisation (nq = 2), with which
// the reference's FD-of-Jacobians mode refuses to run (nq == nv required, problem.hpp:78-81).  Follows the URDF semantics
// Pinocchio's parser implements:
//   * the root link (never a child) is attached to the world; its own inertia does not move and is dropped;
//   * <joint type="fixed"> merges the child link into the body of the supporting joint: the joint's <origin> composes
//     into the placement of everything below it, the child's inertia is added to the body (parallel-axis theorem), and
//     the child link becomes an operational frame of that body;
//   * <origin xyz rpy>: translation, then fixed-axis roll-pitch-yaw, R = Rz(yaw) Ry(pitch) Rx(roll);
//   * <inertial><origin> places the centre of mass and rotates the inertia tensor into the link frame;
//   * every link is a BODY frame (name = link name) on its supporting joint, in document order of the traversal.
// Anything else (floating / planar joints, a link with two movable sub-trees, more joints or frames than
// the kernels are built for) is refused with DDPB_ERR_UNSUPPORTED; malformed text with DDPB_ERR_INVALID.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ddpb.h"

namespace ddpb {
namespace urdf {

// ---- a minimal XML reader: elements, attributes, comments, declarations; text content is ignored ----------------------
struct Node {
  std::string name;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<Node>> kids;
  const Node* child(const char* n) const {
    for (auto& k : kids)
      if (k->name == n) return k.get();
    return nullptr;
  }
};

class Reader {
 public:
  Reader(const char* s, size_t n) : p_(s), e_(s + n) {}
  // returns the root element or nullptr (err set)
  std::unique_ptr<Node> parse(std::string& err) {
    skip_misc();
    auto root = element(err);
    if (!root && err.empty()) err = "no root element";
    return root;
  }

 private:
  const char *p_, *e_;
  bool starts(const char* t) const { const size_t n = std::strlen(t); return (size_t)(e_ - p_) >= n && std::memcmp(p_, t, n) == 0; }
  void skip_ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
  bool skip_until(const char* t) {
    const size_t n = std::strlen(t);
    while ((size_t)(e_ - p_) >= n) {
      if (std::memcmp(p_, t, n) == 0) { p_ += n; return true; }
      ++p_;
    }
    p_ = e_;
    return false;
  }
  // whitespace, text, comments, processing instructions, doctype
  void skip_misc() {
    for (;;) {
      while (p_ < e_ && *p_ != '<') ++p_;
      if (starts("<!--")) { if (!skip_until("-->")) return; }
      else if (starts("<?")) { if (!skip_until("?>")) return; }
      else if (starts("<!")) { if (!skip_until(">")) return; }
      else return;
    }
  }
  static bool name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }
  std::unique_ptr<Node> element(std::string& err) {
    if (p_ >= e_ || *p_ != '<') return nullptr;
    ++p_;
    std::unique_ptr<Node> nd(new Node());
    while (p_ < e_ && name_char(*p_)) nd->name.push_back(*p_++);
    if (nd->name.empty()) { err = "malformed element"; return nullptr; }
    for (;;) {
      skip_ws();
      if (p_ >= e_) { err = "unterminated element <" + nd->name + ">"; return nullptr; }
      if (starts("/>")) { p_ += 2; return nd; }
      if (*p_ == '>') { ++p_; break; }
      std::string key;
      while (p_ < e_ && name_char(*p_)) key.push_back(*p_++);
      skip_ws();
      if (key.empty() || p_ >= e_ || *p_ != '=') { err = "malformed attribute in <" + nd->name + ">"; return nullptr; }
      ++p_;
      skip_ws();
      if (p_ >= e_ || (*p_ != '"' && *p_ != '\'')) { err = "unquoted attribute in <" + nd->name + ">"; return nullptr; }
      const char q = *p_++;
      std::string val;
      while (p_ < e_ && *p_ != q) val.push_back(*p_++);
      if (p_ >= e_) { err = "unterminated attribute in <" + nd->name + ">"; return nullptr; }
      ++p_;
      nd->attr[key] = val;
    }
    for (;;) {  // children until the matching end tag
      skip_misc();
      if (p_ >= e_) { err = "missing </" + nd->name + ">"; return nullptr; }
      if (starts("</")) {
        p_ += 2;
        std::string close;
        while (p_ < e_ && name_char(*p_)) close.push_back(*p_++);
        skip_ws();
        if (close != nd->name || p_ >= e_ || *p_ != '>') { err = "mismatched </" + close + "> for <" + nd->name + ">"; return nullptr; }
        ++p_;
        return nd;
      }
      auto kid = element(err);
      if (!kid) return nullptr;
      nd->kids.push_back(std::move(kid));
    }
  }
};

// ---- small rigid-transform / inertia algebra (row-major 3x3) --------------------------------------------------------------
struct Xf {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double p[3] = {0, 0, 0};
};
inline void mat_mul(const double* A, const double* B, double* C) {
  double t[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) t[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
  std::memcpy(C, t, sizeof(t));
}
inline void mat_vec(const double* A, const double* v, double* o) {
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = A[r * 3] * v[0] + A[r * 3 + 1] * v[1] + A[r * 3 + 2] * v[2];
  std::memcpy(o, t, sizeof(t));
}
inline Xf compose(const Xf& a, const Xf& b) {  // a o b: x -> a.R (b.R x + b.p) + a.p
  Xf c;
  mat_mul(a.R, b.R, c.R);
  mat_vec(a.R, b.p, c.p);
  for (int k = 0; k < 3; ++k) c.p[k] += a.p[k];
  return c;
}
inline void rpy_to_R(const double* rpy, double* R) {  // R = Rz(yaw) Ry(pitch) Rx(roll)
  const double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]), cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
  const double Rx[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr}, Ry[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp}, Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
  double t[9];
  mat_mul(Ry, Rx, t);
  mat_mul(Rz, t, R);
}
// rigid body: mass, centre of mass, inertia about the centre of mass (full symmetric 3x3), all in one frame
struct Body {
  double m = 0.0, c[3] = {0, 0, 0}, I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};
inline Body transformed(const Body& b, const Xf& x) {  // express in the frame x maps into
  Body o;
  o.m = b.m;
  mat_vec(x.R, b.c, o.c);
  for (int k = 0; k < 3; ++k) o.c[k] += x.p[k];
  double t[9], Rt[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rt[r * 3 + c] = x.R[c * 3 + r];
  mat_mul(x.R, b.I, t);
  mat_mul(t, Rt, o.I);
  return o;
}
inline void add_body(Body& a, const Body& b) {  // a <- a + b (composite rigid body, parallel-axis theorem)
  if (b.m == 0.0) return;
  if (a.m == 0.0) { a = b; return; }  // a single part keeps its numbers to the bit
  const double m = a.m + b.m;
  double c[3];
  for (int k = 0; k < 3; ++k) c[k] = m > 0 ? (a.m * a.c[k] + b.m * b.c[k]) / m : 0.0;
  double I[9];
  for (int k = 0; k < 9; ++k) I[k] = a.I[k] + b.I[k];
  const Body* parts[2] = {&a, &b};
  for (const Body* q : parts) {
    const double d[3] = {q->c[0] - c[0], q->c[1] - c[1], q->c[2] - c[2]};
    const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) I[r * 3 + cc] += q->m * ((r == cc ? dd : 0.0) - d[r] * d[cc]);
  }
  a.m = m;
  std::memcpy(a.c, c, sizeof(c));
  std::memcpy(a.I, I, sizeof(I));
}

// ---- the converter ------------------------------------------------------------------------------------------------------------
struct Model {  // owns the arrays a ddpb_model_desc points into
  std::string name;
  std::vector<int> parent, frame_parent, joint_type;
  std::vector<double> placement_rot, placement_trans, axis, mass, com, inertia, frame_rot, frame_trans;
  std::vector<std::string> joint_names, frame_names;
  double gravity[3] = {0.0, 0.0, -9.81};
  ddpb_model_desc desc() const {
    ddpb_model_desc d;
    std::memset(&d, 0, sizeof(d));
    d.kind = DDPB_MODEL_REVOLUTE_CHAIN;
    d.n_joints = (int)parent.size();
    d.parent = parent.data(); d.placement_rot = placement_rot.data(); d.placement_trans = placement_trans.data();
    d.axis = axis.data(); d.mass = mass.data(); d.com = com.data(); d.inertia = inertia.data();
    for (int k = 0; k < 3; ++k) d.gravity[k] = gravity[k];
    d.n_frames = (int)frame_parent.size();
    d.frame_parent = frame_parent.data(); d.frame_rot = frame_rot.data(); d.frame_trans = frame_trans.data();
    d.pend_mass = 1.0; d.pend_length = 1.0;
    d.joint_type = joint_type.empty() ? nullptr : joint_type.data();
    return d;
  }
};

inline bool parse_vec(const std::map<std::string, std::string>& a, const char* key, int n, double* out, std::string& err) {
  auto it = a.find(key);
  if (it == a.end()) return true;  // keep the default
  const char* s = it->second.c_str();
  for (int k = 0; k < n; ++k) {
    char* end = nullptr;
    out[k] = std::strtod(s, &end);
    if (end == s) { err = std::string("attribute ") + key + "=\"" + it->second + "\" needs " + std::to_string(n) + " numbers"; return false; }
    s = end;
  }
  return true;
}
inline bool parse_origin(const Node* n, Xf& x, std::string& err) {
  x = Xf();
  if (!n) return true;
  double rpy[3] = {0, 0, 0};
  if (!parse_vec(n->attr, "xyz", 3, x.p, err) || !parse_vec(n->attr, "rpy", 3, rpy, err)) return false;
  rpy_to_R(rpy, x.R);
  return true;
}
inline bool parse_inertial(const Node* link, Body& b, std::string& err) {
  b = Body();
  const Node* in = link->child("inertial");
  if (!in) return true;
  Xf o;
  if (!parse_origin(in->child("origin"), o, err)) return false;
  Body raw;
  if (const Node* m = in->child("mass")) { if (!parse_vec(m->attr, "value", 1, &raw.m, err)) return false; }
  if (const Node* I = in->child("inertia")) {
    double v[6] = {0, 0, 0, 0, 0, 0};
    const char* keys[6] = {"ixx", "ixy", "ixz", "iyy", "iyz", "izz"};
    for (int k = 0; k < 6; ++k)
      if (!parse_vec(I->attr, keys[k], 1, v + k, err)) return false;
    const double M[9] = {v[0], v[1], v[2], v[1], v[3], v[4], v[2], v[4], v[5]};
    std::memcpy(raw.I, M, sizeof(M));
  }
  b = transformed(raw, o);  // centre of mass and tensor in the link frame
  return true;
}

// returns DDPB_OK or an error code with a message in `err`
inline int model_from_urdf(const char* text, size_t len, Model& out, std::string& err) {
  if (!text || len == 0) { err = "empty URDF text"; return DDPB_ERR_INVALID; }
  Reader rd(text, len);
  std::unique_ptr<Node> root = rd.parse(err);
  if (!root) return DDPB_ERR_INVALID;
  if (root->name != "robot") { err = "root element is <" + root->name + ">, expected <robot>"; return DDPB_ERR_INVALID; }
  out = Model();
  if (root->attr.count("name")) out.name = root->attr["name"];
  std::map<std::string, const Node*> links;
  std::vector<const Node*> joints;
  for (auto& k : root->kids) {
    if (k->name == "link") {
      if (!k->attr.count("name")) { err = "<link> without a name"; return DDPB_ERR_INVALID; }
      links[k->attr["name"]] = k.get();
    } else if (k->name == "joint") {
      joints.push_back(k.get());
    }
  }
  if (links.empty()) { err = "no <link> in the robot"; return DDPB_ERR_INVALID; }
  std::map<std::string, std::vector<const Node*>> by_parent;
  std::map<std::string, int> is_child;
  for (const Node* j : joints) {
    const Node *pa = j->child("parent"), *ch = j->child("child");
    if (!j->attr.count("name") || !j->attr.count("type") || !pa || !ch || !pa->attr.count("link") || !ch->attr.count("link")) {
      err = "<joint> needs name, type, <parent link> and <child link>"; return DDPB_ERR_INVALID;
    }
    const std::string pl = pa->attr.at("link"), cl = ch->attr.at("link");
    if (!links.count(pl) || !links.count(cl)) { err = "joint " + j->attr.at("name") + " refers to an unknown link"; return DDPB_ERR_INVALID; }
    if (is_child[cl]++) { err = "link " + cl + " has two parent joints"; return DDPB_ERR_INVALID; }
    by_parent[pl].push_back(j);
  }
  std::string root_link;
  for (auto& kv : links)
    if (!is_child.count(kv.first) || is_child[kv.first] == 0) {
      if (!root_link.empty()) { err = "two root links: " + root_link + ", " + kv.first; return DDPB_ERR_INVALID; }
      root_link = kv.first;
    }
  if (root_link.empty()) { err = "the links form a cycle"; return DDPB_ERR_INVALID; }

  // depth-first walk.  `jid` = supporting joint of the current link (-1: the world), `x` = placement of the link frame in
  // the frame of that joint (world frame for jid = -1).
  std::vector<Body> bodies;  // per movable joint, in its own joint frame
  struct Item { std::string link; int jid; Xf x; };
  std::vector<Item> stack;
  stack.push_back({root_link, -1, Xf()});
  int visited = 0;
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    if (++visited > (int)links.size()) { err = "the links form a cycle"; return DDPB_ERR_INVALID; }
    const Node* ln = links[it.link];
    Body b;
    if (!parse_inertial(ln, b, err)) return DDPB_ERR_INVALID;
    if (it.jid >= 0) {
      add_body(bodies[it.jid], transformed(b, it.x));
      out.frame_parent.push_back(it.jid);  // BODY frame of the link
      out.frame_rot.insert(out.frame_rot.end(), it.x.R, it.x.R + 9);
      out.frame_trans.insert(out.frame_trans.end(), it.x.p, it.x.p + 3);
      out.frame_names.push_back(it.link);
    }
    int movable = 0;
    auto& kids = by_parent[it.link];
    for (auto rit = kids.rbegin(); rit != kids.rend(); ++rit) {  // reversed: the stack pops them in document order
      const Node* j = *rit;
      const std::string type = j->attr.at("type"), cl = j->child("child")->attr.at("link");
      Xf o;
      if (!parse_origin(j->child("origin"), o, err)) return DDPB_ERR_INVALID;
      const Xf xj = compose(it.x, o);  // joint frame in the supporting joint's frame
      if (type == "fixed") {
        stack.push_back({cl, it.jid, xj});
      } else if (type == "revolute" || type == "continuous" || type == "prismatic") {
        if (++movable > 1) {
          err = "link " + it.link + " carries two movable joints: only serial chains are supported"; return DDPB_ERR_UNSUPPORTED;
        }
        double ax[3] = {1, 0, 0};
        if (const Node* a = j->child("axis")) { if (!parse_vec(a->attr, "xyz", 3, ax, err)) return DDPB_ERR_INVALID; }
        const double nrm = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        if (!(nrm > 0)) { err = "joint " + j->attr.at("name") + " has a zero axis"; return DDPB_ERR_INVALID; }
        const int id = (int)out.parent.size();
        if (it.jid != id - 1) { err = "joint " + j->attr.at("name") + " branches off the chain: only serial chains are supported"; return DDPB_ERR_UNSUPPORTED; }
        out.parent.push_back(it.jid);
        out.placement_rot.insert(out.placement_rot.end(), xj.R, xj.R + 9);
        out.placement_trans.insert(out.placement_trans.end(), xj.p, xj.p + 3);
        for (int k = 0; k < 3; ++k) out.axis.push_back(ax[k] / nrm);
        out.joint_names.push_back(j->attr.at("name"));
        out.joint_type.push_back(type == "prismatic" ? DDPB_JOINT_PRISMATIC : DDPB_JOINT_REVOLUTE);
        bodies.emplace_back();
        stack.push_back({cl, id, Xf()});
      } else {
        err = "joint " + j->attr.at("name") + " has type \"" + type + "\": only revolute, continuous, prismatic and fixed joints are supported";
        return DDPB_ERR_UNSUPPORTED;
      }
    }
  }
  const int n = (int)out.parent.size();
  if (n < 1) { err = "the robot has no movable joint"; return DDPB_ERR_UNSUPPORTED; }
  for (int i = 0; i < n; ++i) {
    const Body& b = bodies[i];
    out.mass.push_back(b.m);
    out.com.insert(out.com.end(), b.c, b.c + 3);
    const double I6[6] = {b.I[0], b.I[1], b.I[2], b.I[4], b.I[5], b.I[8]};
    out.inertia.insert(out.inertia.end(), I6, I6 + 6);
  }
  return DDPB_OK;
}

}  // namespace urdf
}  // namespace ddpb

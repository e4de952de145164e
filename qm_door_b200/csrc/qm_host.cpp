// Host-side ingestion of the reference's input files behind the C-ABI (no GPU involved):
//   URDF  -> qmb200_model_desc    replaces centroidal_model::createPinocchioInterface(urdf, jointNames)
//                                 (qm_interface/src/QMInterface.cpp:408-416; joint list ModelSettings.h:32-38)
//   INFO  -> qmb200_problem_desc / qmb200_solver_desc
//                                 replaces the ocs2::loadData calls of QMInterface.cpp:65-73,85,152-156,199-234,291,306,395-397
//   gait  -> mode sequence templates and tiled mode schedules (QMInterface.cpp:455-480, config/gait.info)
// Boost.PropertyTree and urdfdom are not available in this image, so both formats are parsed here.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/qmb200.h"
#include "qm_mpc.h"

namespace {

thread_local std::string g_host_err;

std::string read_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::invalid_argument("file not found: " + path);   // QMInterface.cpp:41-62 throws std::invalid_argument
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// ---------------------------------------------------------------------------------------- INFO (Boost.PropertyTree)
struct InfoNode {
  std::string value;
  std::vector<std::pair<std::string, std::shared_ptr<InfoNode>>> children;
  const InfoNode* find(const std::string& key) const {
    for (auto& c : children)
      if (c.first == key) return c.second.get();
    return nullptr;
  }
  const InfoNode& at(const std::string& dotted) const {
    const InfoNode* n = this;
    size_t pos = 0;
    while (pos <= dotted.size()) {
      size_t dot = dotted.find('.', pos);
      std::string k = dotted.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
      n = n->find(k);
      if (!n) throw std::runtime_error("INFO key not found: " + dotted);
      if (dot == std::string::npos) break;
      pos = dot + 1;
    }
    return *n;
  }
  double num(const std::string& dotted) const { return std::strtod(at(dotted).value.c_str(), nullptr); }
};

std::vector<std::string> info_tokens(const std::string& text) {
  std::vector<std::string> toks;
  std::istringstream in(text);
  std::string line;
  while (std::getline(in, line)) {
    size_t sc = line.find(';');
    if (sc != std::string::npos) line = line.substr(0, sc);
    size_t i = 0;
    while (i < line.size()) {
      char c = line[i];
      if (isspace((unsigned char)c)) { ++i; continue; }
      if (c == '{' || c == '}') { toks.push_back(std::string(1, c)); ++i; continue; }
      if (c == '"') {
        size_t j = line.find('"', i + 1);
        if (j == std::string::npos) j = line.size();
        toks.push_back(line.substr(i + 1, j - i - 1));
        i = j + 1;
        continue;
      }
      size_t j = i;
      while (j < line.size() && !isspace((unsigned char)line[j]) && line[j] != '{' && line[j] != '}') ++j;
      toks.push_back(line.substr(i, j - i));
      i = j;
    }
    toks.push_back("\n");
  }
  return toks;
}

void info_block(const std::vector<std::string>& t, size_t& pos, InfoNode& out) {
  while (pos < t.size()) {
    if (t[pos] == "\n") { ++pos; continue; }
    if (t[pos] == "}") { ++pos; return; }
    auto node = std::make_shared<InfoNode>();
    std::string key = t[pos++];
    if (pos < t.size() && t[pos] != "\n" && t[pos] != "{" && t[pos] != "}") node->value = t[pos++];
    while (pos < t.size() && t[pos] == "\n") ++pos;
    if (pos < t.size() && t[pos] == "{") { ++pos; info_block(t, pos, *node); }
    out.children.emplace_back(key, node);
  }
}

InfoNode parse_info(const std::string& path) {
  InfoNode root;
  auto toks = info_tokens(read_file(path));
  size_t pos = 0;
  info_block(toks, pos, root);
  return root;
}

// ocs2::loadData::loadEigenMatrix: "(i,j) value" entries, optional "scaling"
void load_matrix(const InfoNode& root, const std::string& name, int rows, int cols, double* out) {
  const InfoNode& n = root.at(name);
  for (int i = 0; i < rows * cols; ++i) out[i] = 0.0;
  double scaling = 1.0;
  if (const InfoNode* s = n.find("scaling")) scaling = std::strtod(s->value.c_str(), nullptr);
  for (auto& c : n.children) {
    int i, j;
    if (sscanf(c.first.c_str(), "(%d,%d)", &i, &j) == 2 && i >= 0 && i < rows && j >= 0 && j < cols)
      out[i * cols + j] = std::strtod(c.second->value.c_str(), nullptr);
  }
  for (int i = 0; i < rows * cols; ++i) out[i] *= scaling;
}

std::vector<std::string> load_list(const InfoNode& n) {
  std::vector<std::pair<int, std::string>> items;
  for (auto& c : n.children) {
    int i;
    if (sscanf(c.first.c_str(), "[%d]", &i) == 1) items.emplace_back(i, c.second->value);
  }
  std::sort(items.begin(), items.end(), [](auto& a, auto& b) { return a.first < b.first; });
  std::vector<std::string> out;
  for (auto& it : items) out.push_back(it.second);
  return out;
}

int mode_from_name(const std::string& s) {   // [upstream] ocs2::legged_robot::string2ModeNumber: LF=8 RF=4 LH=2 RH=1
  if (s == "STANCE") return 15;
  if (s == "FLY") return 0;
  int m = 0;
  std::istringstream in(s);
  std::string tok;
  while (std::getline(in, tok, '_')) {
    if (tok == "LF") m |= 8;
    else if (tok == "RF") m |= 4;
    else if (tok == "LH") m |= 2;
    else if (tok == "RH") m |= 1;
    else throw std::runtime_error("unknown mode name: " + s);
  }
  return m;
}

// ---------------------------------------------------------------------------------------- minimal XML (URDF subset)
struct XmlNode {
  std::string name;
  std::map<std::string, std::string> attr;
  std::vector<std::shared_ptr<XmlNode>> children;
  const XmlNode* child(const std::string& n) const {
    for (auto& c : children)
      if (c->name == n) return c.get();
    return nullptr;
  }
  std::string get(const std::string& k, const std::string& def = "") const {
    auto it = attr.find(k);
    return it == attr.end() ? def : it->second;
  }
};

struct XmlParser {
  const std::string& s;
  size_t p = 0;
  explicit XmlParser(const std::string& text) : s(text) {}
  void skip_ws() { while (p < s.size() && isspace((unsigned char)s[p])) ++p; }
  bool starts(const char* lit) const { return s.compare(p, strlen(lit), lit) == 0; }
  void skip_misc() {
    for (;;) {
      skip_ws();
      if (starts("<?")) { p = s.find("?>", p); p = (p == std::string::npos) ? s.size() : p + 2; }
      else if (starts("<!--")) { p = s.find("-->", p); p = (p == std::string::npos) ? s.size() : p + 3; }
      else if (starts("<!")) { p = s.find('>', p); p = (p == std::string::npos) ? s.size() : p + 1; }
      else break;
    }
  }
  std::shared_ptr<XmlNode> element() {
    skip_misc();
    if (p >= s.size() || s[p] != '<') throw std::runtime_error("URDF: expected '<'");
    ++p;
    auto n = std::make_shared<XmlNode>();
    size_t b = p;
    while (p < s.size() && !isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/') ++p;
    n->name = s.substr(b, p - b);
    for (;;) {
      skip_ws();
      if (p >= s.size()) throw std::runtime_error("URDF: unterminated tag");
      if (s[p] == '/') { p += 2; return n; }
      if (s[p] == '>') { ++p; break; }
      size_t kb = p;
      while (p < s.size() && s[p] != '=' && !isspace((unsigned char)s[p])) ++p;
      std::string key = s.substr(kb, p - kb);
      skip_ws();
      if (s[p] != '=') throw std::runtime_error("URDF: attribute without value");
      ++p;
      skip_ws();
      char q = s[p++];
      size_t vb = p;
      while (p < s.size() && s[p] != q) ++p;
      n->attr[key] = s.substr(vb, p - vb);
      ++p;
    }
    for (;;) {
      // text content is irrelevant for URDF
      while (p < s.size() && s[p] != '<') ++p;
      skip_misc();
      if (p >= s.size()) throw std::runtime_error("URDF: unterminated element " + n->name);
      if (starts("</")) { p = s.find('>', p); p = (p == std::string::npos) ? s.size() : p + 1; return n; }
      n->children.push_back(element());
    }
  }
};

struct V3 { double v[3]; };
struct M3 { double m[9]; };
M3 eye3() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
M3 mul(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return c;
}
M3 transpose(const M3& a) { return M3{{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
V3 mul(const M3& a, const V3& b) {
  V3 c;
  for (int i = 0; i < 3; ++i) c.v[i] = a.m[3 * i] * b.v[0] + a.m[3 * i + 1] * b.v[1] + a.m[3 * i + 2] * b.v[2];
  return c;
}
V3 add(const V3& a, const V3& b) { return V3{{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
V3 parse_v3(const std::string& s) {
  V3 v{{0, 0, 0}};
  if (sscanf(s.c_str(), "%lf %lf %lf", &v.v[0], &v.v[1], &v.v[2]) != 3) throw std::runtime_error("URDF: bad vector '" + s + "'");
  return v;
}
M3 rpy(const V3& a) {
  const double cr = cos(a.v[0]), sr = sin(a.v[0]), cp = cos(a.v[1]), sp = sin(a.v[1]), cy = cos(a.v[2]), sy = sin(a.v[2]);
  M3 Rx{{1, 0, 0, 0, cr, -sr, 0, sr, cr}}, Ry{{cp, 0, sp, 0, 1, 0, -sp, 0, cp}}, Rz{{cy, -sy, 0, sy, cy, 0, 0, 0, 1}};
  return mul(Rz, mul(Ry, Rx));
}
void origin_of(const XmlNode* el, M3& R, V3& p) {
  R = eye3(); p = V3{{0, 0, 0}};
  const XmlNode* o = el ? el->child("origin") : nullptr;
  if (!o) return;
  R = rpy(parse_v3(o->get("rpy", "0 0 0")));
  p = parse_v3(o->get("xyz", "0 0 0"));
}

const char* kJointNames[18] = {"LF_HAA", "LF_HFE", "LF_KFE", "RF_HAA", "RF_HFE", "RF_KFE", "LH_HAA", "LH_HFE", "LH_KFE",
                               "RH_HAA", "RH_HFE", "RH_KFE", "z1_joint_1", "z1_joint_2", "z1_joint_3", "z1_joint_4",
                               "z1_joint_5", "z1_joint_6"};                      // ModelSettings.h:32-36
const char* kContactNames[4] = {"LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"};      // ModelSettings.h:38
const char* kEeFrame = "z1_end_effector";                                          // task.info:20

struct BodyPart { double m; V3 c; M3 I; };
struct Builder {
  std::map<std::string, const XmlNode*> links;
  std::map<std::string, std::vector<const XmlNode*>> joints_by_parent;
  qmb200_model_desc* M;
  int nj = 0;
  std::vector<std::vector<BodyPart>> parts;
  std::map<std::string, std::pair<int, std::pair<V3, M3>>> frames;

  int add_joint(int parent, int type, const V3& axis, const M3& R, const V3& p, double lo, double hi, double eff) {
    if (nj >= QM_NJ) throw std::runtime_error("URDF: more than 24 degrees of freedom");
    const int j = nj++;
    M->parent[j] = parent; M->jtype[j] = type;
    for (int k = 0; k < 3; ++k) { M->axis[j][k] = axis.v[k]; M->pp[j][k] = p.v[k]; }
    for (int k = 0; k < 9; ++k) M->Rp[j][k] = R.m[k];
    M->lower[j] = lo; M->upper[j] = hi; M->effort[j] = eff;
    parts.emplace_back();
    return j;
  }
  void visit(const std::string& link, int jidx, const M3& R, const V3& p) {
    const XmlNode* L = links.at(link);
    frames[link] = {jidx, {p, R}};
    if (const XmlNode* in = L->child("inertial")) {
      M3 Ro; V3 po;
      origin_of(in, Ro, po);
      const XmlNode* mass = in->child("mass");
      const XmlNode* I = in->child("inertia");
      if (!mass || !I) throw std::runtime_error("URDF: inertial without mass/inertia in link " + link);
      auto g = [&](const char* k) { return std::strtod(I->get(k, "0").c_str(), nullptr); };
      M3 Ic{{g("ixx"), g("ixy"), g("ixz"), g("ixy"), g("iyy"), g("iyz"), g("ixz"), g("iyz"), g("izz")}};
      M3 Rb = mul(R, Ro);
      parts[jidx].push_back(BodyPart{std::strtod(mass->get("value").c_str(), nullptr), add(p, mul(R, po)), mul(Rb, mul(Ic, transpose(Rb)))});
    }
    auto it = joints_by_parent.find(link);
    if (it == joints_by_parent.end()) return;
    std::vector<const XmlNode*> js = it->second;
    // urdfdom keeps child links in a name-sorted map -> Pinocchio's joint order [upstream]
    std::sort(js.begin(), js.end(), [](const XmlNode* a, const XmlNode* b) { return a->child("child")->get("link") < b->child("child")->get("link"); });
    for (const XmlNode* j : js) {
      M3 Rj; V3 pj;
      origin_of(j, Rj, pj);
      M3 Rc = mul(R, Rj);
      V3 pc = add(p, mul(R, pj));
      const std::string child = j->child("child")->get("link");
      const std::string type = j->get("type");
      bool actuated = false;
      for (auto n : kJointNames) actuated |= (j->get("name") == n);
      if (type == "fixed" || !actuated) {
        visit(child, jidx, Rc, pc);   // joints outside the list are locked at their neutral position (q = 0)
      } else {
        if (type != "revolute" && type != "continuous") throw std::runtime_error("URDF: unsupported joint type " + type);
        const XmlNode* ax = j->child("axis");
        const XmlNode* lim = j->child("limit");
        auto lg = [&](const char* k, double d) { return lim && !lim->get(k).empty() ? std::strtod(lim->get(k).c_str(), nullptr) : d; };
        const int nw = add_joint(jidx, 1, ax ? parse_v3(ax->get("xyz", "1 0 0")) : V3{{1, 0, 0}}, Rc, pc, lg("lower", -1e30), lg("upper", 1e30), lg("effort", 0));
        visit(child, nw, eye3(), V3{{0, 0, 0}});
      }
    }
  }
};

void finalize_model(qmb200_model_desc* M) {
  for (int j = 0; j < QM_NJ; ++j) M->depth[j] = M->parent[j] < 0 ? 0 : M->depth[M->parent[j]] + 1;
  M->max_depth = 0;
  for (int j = 0; j < QM_NJ; ++j) {
    M->max_depth = std::max(M->max_depth, M->depth[j]);
    uint32_t pm = 0;
    for (int k = j; k >= 0; k = M->parent[k]) pm |= 1u << k;
    M->pathmask[j] = pm;
  }
  for (int j = 0; j < QM_NJ; ++j) {
    uint32_t sm = 0;
    for (int i = 0; i < QM_NJ; ++i)
      if ((M->pathmask[i] >> j) & 1u) sm |= 1u << i;
    M->submask[j] = sm;
  }
  M->total_mass = 0;
  for (int j = 0; j < QM_NJ; ++j) M->total_mass += M->mass[j];
  // standard floating base? (lets the kinematics place joints 0..5 in closed form instead of walking the chain)
  static const int kType[6] = {0, 0, 0, 1, 1, 1};
  static const double kAxis[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
  bool stdroot = true;
  for (int k = 0; k < 6; ++k) {
    stdroot = stdroot && M->parent[k] == k - 1 && M->jtype[k] == kType[k];
    for (int c = 0; c < 3; ++c) stdroot = stdroot && M->axis[k][c] == kAxis[k][c] && M->pp[k][c] == 0.0;
    for (int c = 0; c < 9; ++c) stdroot = stdroot && M->Rp[k][c] == ((c % 4 == 0) ? 1.0 : 0.0);
  }
  M->root6_standard = stdroot ? 1 : 0;
}

template <class F>
int guarded(F f) {
  try { f(); return 0; }
  catch (const std::exception& e) { g_host_err = e.what(); return -2; }
}

}  // namespace

extern "C" {

const char* qmb200_last_error(void) { return g_host_err.c_str(); }
void qmb200_set_error_(const char* msg) { g_host_err = msg ? msg : ""; }

int qmb200_load_urdf(const char* urdf_path, qmb200_model_desc* model) {
  return guarded([&]() {
    if (!urdf_path || !model) throw std::invalid_argument("qmb200_load_urdf: null argument");
    std::string text = read_file(urdf_path);
    XmlParser xp(text);
    auto root = xp.element();
    if (root->name != "robot") throw std::runtime_error("URDF: root element is not <robot>");
    memset(model, 0, sizeof(*model));
    Builder b;
    b.M = model;
    std::map<std::string, bool> is_child;
    for (auto& c : root->children) {
      if (c->name == "link") b.links[c->get("name")] = c.get();
      if (c->name == "joint") {
        if (!c->child("parent") || !c->child("child")) throw std::runtime_error("URDF: joint without parent/child");
        b.joints_by_parent[c->child("parent")->get("link")].push_back(c.get());
        is_child[c->child("child")->get("link")] = true;
      }
    }
    std::string base;
    for (auto& l : b.links)
      if (!is_child.count(l.first)) { if (!base.empty()) throw std::runtime_error("URDF: several root links"); base = l.first; }
    if (base.empty()) throw std::runtime_error("URDF: no root link");
    // floating base = composite (translation xyz, spherical ZYX with Euler-rate velocities) as six one-DoF joints [upstream]
    const V3 ex{{1, 0, 0}}, ey{{0, 1, 0}}, ez{{0, 0, 1}}, z{{0, 0, 0}};
    b.add_joint(-1, 0, ex, eye3(), z, -1e30, 1e30, 0);
    b.add_joint(0, 0, ey, eye3(), z, -1e30, 1e30, 0);
    b.add_joint(1, 0, ez, eye3(), z, -1e30, 1e30, 0);
    b.add_joint(2, 1, ez, eye3(), z, -1e30, 1e30, 0);
    b.add_joint(3, 1, ey, eye3(), z, -1e30, 1e30, 0);
    b.add_joint(4, 1, ex, eye3(), z, -1e30, 1e30, 0);
    b.visit(base, 5, eye3(), z);
    if (b.nj != QM_NJ) throw std::runtime_error("URDF: expected 18 actuated joints, found " + std::to_string(b.nj - 6));
    model->nj = QM_NJ;
    for (int j = 0; j < QM_NJ; ++j) {
      double mt = 0;
      V3 c{{0, 0, 0}};
      for (auto& pt : b.parts[j]) { mt += pt.m; for (int k = 0; k < 3; ++k) c.v[k] += pt.m * pt.c.v[k]; }
      if (mt == 0) continue;
      for (int k = 0; k < 3; ++k) c.v[k] /= mt;
      double I[9] = {0};
      for (auto& pt : b.parts[j]) {
        const double d[3] = {pt.c.v[0] - c.v[0], pt.c.v[1] - c.v[1], pt.c.v[2] - c.v[2]};
        const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        for (int r = 0; r < 3; ++r)
          for (int cc = 0; cc < 3; ++cc) I[3 * r + cc] += pt.I.m[3 * r + cc] + pt.m * ((r == cc ? dd : 0.0) - d[r] * d[cc]);
      }
      model->mass[j] = mt;
      for (int k = 0; k < 3; ++k) model->com[j][k] = c.v[k];
      for (int k = 0; k < 9; ++k) model->inertia[j][k] = I[k];
    }
    for (int f = 0; f < 4; ++f) {
      auto it = b.frames.find(kContactNames[f]);
      if (it == b.frames.end()) throw std::runtime_error(std::string("URDF: contact frame missing: ") + kContactNames[f]);
      model->foot_joint[f] = it->second.first;
      for (int k = 0; k < 3; ++k) model->foot_off[f][k] = it->second.second.first.v[k];
    }
    auto ee = b.frames.find(kEeFrame);
    if (ee == b.frames.end()) throw std::runtime_error("URDF: end-effector frame missing");
    model->ee_joint = ee->second.first;
    for (int k = 0; k < 3; ++k) model->ee_off[k] = ee->second.second.first.v[k];
    for (int k = 0; k < 9; ++k) model->ee_Roff[k] = ee->second.second.second.m[k];
    finalize_model(model);
  });
}

int qmb200_load_targets(const char* task_info, const char* reference_info, qmb200_target_desc* D) {
  return guarded([&]() {
    if (!task_info || !reference_info || !D) throw std::invalid_argument("qmb200_load_targets: null argument");
    InfoNode t = parse_info(task_info);
    InfoNode r = parse_info(reference_info);
    memset(D, 0, sizeof(*D));
    // QmTargetTrajectoriesPublisher_node.cpp:268-272
    D->com_height = r.num("comHeight");
    load_matrix(r, "defaultJointState", 18, 1, D->default_joint_state);
    D->target_rotation_velocity = r.num("targetRotationVelocity");
    D->target_displacement_velocity = r.num("targetDisplacementVelocity");
    D->time_to_target = t.num("mpc.timeHorizon");
    D->arm_dist = 0.6;         // qm_controllers/include/qm_controllers/StartingPosition.h:13
    D->feet_height = 0.0;      // runtime value (feetHeightCallback, :28-35)
  });
}

int qmb200_load_problem(const char* task_info, const char* reference_info, const qmb200_model_desc* M,
                        qmb200_problem_desc* P, qmb200_solver_desc* S, double* x_init) {
  return guarded([&]() {
    if (!task_info || !M || !P || !S) throw std::invalid_argument("qmb200_load_problem: null argument");
    InfoNode t = parse_info(task_info);
    memset(P, 0, sizeof(*P));
    memset(S, 0, sizeof(*S));
    double xi[30];
    load_matrix(t, "initialState", 30, 1, xi);
    if (x_init) memcpy(x_init, xi, sizeof(xi));
    load_matrix(t, "Q", 30, 30, P->Q);
    std::vector<double> Rt(900);
    load_matrix(t, "R", 30, 30, Rt.data());
    // QMInterface::initializeInputCostWeight (QMInterface.cpp:274-299): leg block <- J' R J at the initial state
    std::vector<double> w(qm::KW_SIZE);
    qm::kin_eval(qm::SerialGroup(), *M, xi, (const double*)nullptr, false, w.data());
    double J[12][12];
    for (int r = 0; r < 12; ++r)
      for (int c = 0; c < 12; ++c) J[r][c] = w[qm::KW_FJ + r * QM_NJ + 6 + c];
    memcpy(P->R, Rt.data(), sizeof(double) * 900);
    for (int a = 0; a < 12; ++a)
      for (int b = 0; b < 12; ++b) {
        double acc = 0;
        for (int r = 0; r < 12; ++r)
          for (int c = 0; c < 12; ++c) acc += J[r][a] * Rt[(12 + r) * 30 + 12 + c] * J[c][b];
        P->R[(12 + a) * 30 + 12 + b] = acc;
      }
    P->mu_ee_pos = t.num("endEffector.muPosition");
    P->mu_ee_ori = t.num("endEffector.muOrientation");
    P->mu_fee_pos = t.num("finalEndEffector.muPosition");
    P->mu_fee_ori = t.num("finalEndEffector.muOrientation");
    P->fric_mu = t.num("frictionConeSoftConstraint.frictionCoefficient");
    P->fric_bar_mu = t.num("frictionConeSoftConstraint.mu");
    P->fric_bar_delta = t.num("frictionConeSoftConstraint.delta");
    P->fric_reg = 25.0; P->fric_grip = 0.0; P->fric_hess_shift = 1e-6;   // [upstream] FrictionConeConstraint::Config defaults
    P->pos_bar_mu = t.num("jointPositionLimits.mu");
    P->pos_bar_delta = t.num("jointPositionLimits.delta");
    P->vel_bar_mu = t.num("jointVelocityLimits.mu");
    P->vel_bar_delta = t.num("jointVelocityLimits.delta");
    load_matrix(t, "jointVelocityLimits.lowerBound.arm", 6, 1, P->arm_vel_lo);
    load_matrix(t, "jointVelocityLimits.upperBound.arm", 6, 1, P->arm_vel_hi);
    for (int i = 0; i < 6; ++i) { P->arm_pos_lo[i] = M->lower[18 + i]; P->arm_pos_hi[i] = M->upper[18 + i]; }
    double off = 0, v, d1, d2;
    for (int i = 0; i < 6; ++i) {   // StateInputSoftBoxConstraint::initializeOffset(0, 0, 0) (QMInterface.cpp:257)
      qm::relaxed_barrier(0.0 - P->arm_pos_lo[i], P->pos_bar_mu, P->pos_bar_delta, &v, &d1, &d2); off += v;
      qm::relaxed_barrier(P->arm_pos_hi[i] - 0.0, P->pos_bar_mu, P->pos_bar_delta, &v, &d1, &d2); off += v;
      qm::relaxed_barrier(0.0 - P->arm_vel_lo[i], P->vel_bar_mu, P->vel_bar_delta, &v, &d1, &d2); off += v;
      qm::relaxed_barrier(P->arm_vel_hi[i] - 0.0, P->vel_bar_mu, P->vel_bar_delta, &v, &d1, &d2); off += v;
    }
    P->box_offset = off;
    P->swing_liftoff_vel = t.num("swing_trajectory_config.liftOffVelocity");
    P->swing_touchdown_vel = t.num("swing_trajectory_config.touchDownVelocity");
    P->swing_height = t.num("swing_trajectory_config.swingHeight");
    P->swing_time_scale = t.num("swing_trajectory_config.swingTimeScale");
    P->gravity = 9.81;
    S->dt = t.num("sqp.dt");
    S->horizon = t.num("mpc.timeHorizon");
    S->delta_tol = t.num("sqp.deltaTol");
    S->g_max = t.num("sqp.g_max");
    S->g_min = t.num("sqp.g_min");
    S->alpha_decay = 0.5; S->alpha_min = 1e-4; S->gamma_c = 1e-6; S->armijo_factor = 1e-4;   // [upstream] sqp::Settings defaults
    S->weak_eps = 1e-6; S->dt_min = 1e-8;
    S->sqp_iterations = (int)std::lround(t.num("sqp.sqpIteration"));   // task.info:80
    S->cost_tol = 1e-4;                                                  // [upstream] sqp::Settings::costTol
    S->max_nodes = (int)std::lround(S->horizon / S->dt) + 1 + 24;
    S->max_events = 32;
    S->max_targets = 2;
    if (reference_info) parse_info(reference_info);   // existence / syntax check (QMInterface.cpp:57-62)
  });
}

int qmb200_load_gait(const char* gait_info, const char* gait_name, int32_t capacity, double* switching_times, int32_t* modes,
                     int32_t* num_modes) {
  return guarded([&]() {
    if (!gait_info || !gait_name || !switching_times || !modes || !num_modes) throw std::invalid_argument("qmb200_load_gait: null argument");
    InfoNode g = parse_info(gait_info);
    const InfoNode& n = g.at(gait_name);
    auto ms = load_list(n.at("modeSequence"));
    auto ts = load_list(n.at("switchingTimes"));
    if (ts.size() != ms.size() + 1) throw std::runtime_error("gait: switchingTimes must have one more entry than modeSequence");
    if ((int)ms.size() > capacity) throw std::runtime_error("gait: capacity too small");
    for (size_t i = 0; i < ms.size(); ++i) modes[i] = mode_from_name(ms[i]);
    for (size_t i = 0; i < ts.size(); ++i) switching_times[i] = std::strtod(ts[i].c_str(), nullptr);
    *num_modes = (int32_t)ms.size();
  });
}

// [upstream] GaitSchedule: STANCE until t_insert, then the template tiled until an event >= t_upper, then STANCE.
int qmb200_tile_schedule(const double* sw, const int32_t* tmodes, int32_t nm, double t_insert, double t_upper, int32_t capacity,
                         double* events, int32_t* mode_sequence, int32_t* num_events) {
  return guarded([&]() {
    if (!sw || !tmodes || !events || !mode_sequence || !num_events || nm < 1) throw std::invalid_argument("qmb200_tile_schedule: bad argument");
    int ne = 0;
    auto push = [&](double e, int mode_before) {
      if (ne >= capacity) throw std::runtime_error("qmb200_tile_schedule: event capacity exceeded");
      mode_sequence[ne] = mode_before;
      events[ne++] = e;
    };
    push(t_insert, 15);
    while (events[ne - 1] < t_upper)
      for (int i = 0; i < nm; ++i) push(events[ne - 1] + (sw[i + 1] - sw[i]), tmodes[i]);
    // mode_sequence[k] = mode of phase k (the phase that ends at event k); the final phase is STANCE
    for (int k = ne; k <= capacity; ++k) mode_sequence[k] = 15;
    for (int k = ne; k < capacity; ++k) events[k] = 1e30;
    *num_events = ne;
  });
}

// WbcBase::loadTasksSetting (qm_wbc/src/WbcBase.cpp:597-627) + default gains of qm_wbc/cfg/wbcWigeht.cfg:7-47
int qmb200_load_wbc(const char* task_info, const qmb200_model_desc* M, qmb200_wbc_desc* C) {
  return guarded([&]() {
    if (!task_info || !M || !C) throw std::invalid_argument("qmb200_load_wbc: null argument");
    InfoNode t = parse_info(task_info);
    memset(C, 0, sizeof(*C));
    C->kp_swing = 350; C->kd_swing = 37;
    C->kp_base_height = 400; C->kd_base_height = 140;
    C->kp_base_linear = 400; C->kd_base_linear = 100;
    C->kp_base_angular = 400; C->kd_base_angular = 140;
    const double kpj[6] = {4000, 4200, 4000, 4000, 4200, 6000};
    for (int i = 0; i < 6; ++i) { C->kp_arm_joint[i] = kpj[i]; C->kd_arm_joint[i] = 75; }
    for (int i = 0; i < 3; ++i) { C->kp_ee_linear[i] = 3000; C->kd_ee_linear[i] = 75; C->kp_ee_angular[i] = 2000; C->kd_ee_angular[i] = 75; }
    C->friction_mu = t.num("frictionConeTask.frictionCoefficient");
    // WbcBase::loadTasksSetting (WbcBase.cpp:599-604): the first leg's (HAA, HFE, KFE) limits are replicated over the four legs
    // (formulateTorqueLimitsTask, :409-410); the arm takes its own six.
    for (int i = 0; i < 12; ++i) C->tau_max[i] = M->effort[6 + i % 3];
    for (int i = 12; i < 18; ++i) C->tau_max[i] = M->effort[6 + i];
    C->swing_weight = 100.0;   // HierarchicalWbc.cpp:29
    C->init_time = 10.0;       // HierarchicalWbc.cpp:32
    C->gravity = 9.81;
    C->mpc_variant = 0;
  });
}

// Defaults of the control law and the simulated actuator: QMController.cpp:181-190 (legs kp 0, kd 3 after 10 s),
// qm_controllers/cfg/weight.cfg:7-8 (arm kp 0, kd 0.5), qm_gazebo/config/default.yaml:2 (delay 0.009 s).
void qmb200_actuator_defaults(qmb200_actuator_desc* D) {
  if (!D) return;
  D->leg_kp = 0.0; D->leg_kd = 3.0;
  D->arm_kp = 0.0; D->arm_kd = 0.5;
  D->leg_enable_time = 10.0;
  D->delay_ns = 9000000;
}

}  // extern "C"

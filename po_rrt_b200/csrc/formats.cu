// formats.cu -- the on-disk formats either side of the hot path, host code only:
//   * PNM gray maps, P2 (ASCII) and P5 (binary), 8 bit: the decode step of Map::open_image / MapShelfDomain (map_io.rs:98-105,
//     map_shelves_io.rs:92-99: image::open(..) must yield ImageLuma8, anything else panics with "Wrong image format!");
//   * PTOGraph JSON (pto_graph.rs:22-118: serde_json of SerializablePTOGraph { nodes: [{state, validity_id, parents: [{id,
//     validity_id}], children: [..]}], validities: [[bool]] }), read into / written from the CSR arrays the value-backup entry
//     points take -- stored (insertion) order of parents and children is kept, it decides tie-breaks downstream.
// A roadmap saved by the Rust planner (`pto_graph::save`) can be loaded here and handed to porrt_sssp_worlds / porrt_belief_vi; a
// roadmap built on the GPU can be written back for `pto_graph::load`.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>

#include "common.cuh"

namespace {

int32_t fmt_fail(porrt_ctx* ctx, int32_t code, const std::string& msg) { return porrt_fail(ctx, code, msg); }

bool read_file(const char* path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}

// ------------------------------------------------------------------------------------------------ PNM
// next header token: skips white space and '#' comments (to the end of the line)
bool pnm_token(const std::string& d, size_t* pos, std::string* tok) {
  size_t p = *pos;
  for (;;) {
    while (p < d.size() && isspace((unsigned char)d[p])) ++p;
    if (p < d.size() && d[p] == '#') { while (p < d.size() && d[p] != '\n') ++p; continue; }
    break;
  }
  const size_t s = p;
  while (p < d.size() && !isspace((unsigned char)d[p])) ++p;
  if (p == s) return false;
  *tok = d.substr(s, p - s);
  *pos = p;
  return true;
}

bool to_int(const std::string& s, long* out) {
  if (s.empty()) return false;
  char* end = nullptr;
  errno = 0;
  const long v = strtol(s.c_str(), &end, 10);
  if (errno || *end) return false;
  *out = v;
  return true;
}

// ------------------------------------------------------------------------------------------------ JSON (reader)
struct JVal {
  enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;
  const JVal* get(const char* key) const {
    for (const auto& kv : obj) if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct JParser {
  const std::string& d;
  size_t p = 0;
  std::string err;
  explicit JParser(const std::string& s) : d(s) {}
  void ws() { while (p < d.size() && (d[p] == ' ' || d[p] == '\n' || d[p] == '\t' || d[p] == '\r')) ++p; }
  bool fail(const char* m) { if (err.empty()) err = std::string(m) + " at byte " + std::to_string(p); return false; }
  bool lit(const char* s) { const size_t n = strlen(s); if (d.compare(p, n, s) != 0) return fail("bad literal"); p += n; return true; }
  bool string(std::string* out) {
    if (p >= d.size() || d[p] != '"') return fail("expected string");
    ++p;
    out->clear();
    while (p < d.size() && d[p] != '"') {
      if (d[p] == '\\') {
        if (p + 1 >= d.size()) return fail("bad escape");
        const char c = d[p + 1];
        p += 2;
        switch (c) {
          case '"': out->push_back('"'); break; case '\\': out->push_back('\\'); break; case '/': out->push_back('/'); break;
          case 'b': out->push_back('\b'); break; case 'f': out->push_back('\f'); break; case 'n': out->push_back('\n'); break;
          case 'r': out->push_back('\r'); break; case 't': out->push_back('\t'); break;
          case 'u': if (p + 4 > d.size()) return fail("bad \\u escape"); out->push_back('?'); p += 4; break;   // keys here are ASCII
          default: return fail("bad escape");
        }
      } else out->push_back(d[p++]);
    }
    if (p >= d.size()) return fail("unterminated string");
    ++p;
    return true;
  }
  bool value(JVal* v, int depth = 0) {
    if (depth > 64) return fail("nesting too deep");
    ws();
    if (p >= d.size()) return fail("unexpected end");
    const char c = d[p];
    if (c == '{') {
      v->kind = JVal::OBJ;
      ++p; ws();
      if (p < d.size() && d[p] == '}') { ++p; return true; }
      for (;;) {
        ws();
        std::string key;
        if (!string(&key)) return false;
        ws();
        if (p >= d.size() || d[p] != ':') return fail("expected ':'");
        ++p;
        v->obj.emplace_back(key, JVal());
        if (!value(&v->obj.back().second, depth + 1)) return false;
        ws();
        if (p < d.size() && d[p] == ',') { ++p; continue; }
        if (p < d.size() && d[p] == '}') { ++p; return true; }
        return fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      v->kind = JVal::ARR;
      ++p; ws();
      if (p < d.size() && d[p] == ']') { ++p; return true; }
      for (;;) {
        v->arr.emplace_back();
        if (!value(&v->arr.back(), depth + 1)) return false;
        ws();
        if (p < d.size() && d[p] == ',') { ++p; continue; }
        if (p < d.size() && d[p] == ']') { ++p; return true; }
        return fail("expected ',' or ']'");
      }
    }
    if (c == '"') { v->kind = JVal::STR; return string(&v->str); }
    if (c == 't') { v->kind = JVal::BOOL; v->b = true; return lit("true"); }
    if (c == 'f') { v->kind = JVal::BOOL; v->b = false; return lit("false"); }
    if (c == 'n') { v->kind = JVal::NUL; return lit("null"); }
    if (c == '-' || (c >= '0' && c <= '9')) {
      char* end = nullptr;
      v->kind = JVal::NUM;
      v->num = strtod(d.c_str() + p, &end);   // correctly rounded (glibc): the f64 the file's digits denote
      if (end == d.c_str() + p) return fail("bad number");
      p = (size_t)(end - d.c_str());
      return true;
    }
    return fail("unexpected character");
  }
};

// a JSON number that must be a usize (serde: "invalid type" otherwise)
bool as_index(const JVal* v, int64_t* out) {
  if (!v || v->kind != JVal::NUM || !(v->num >= 0.0) || v->num > 9.0e15 || v->num != std::floor(v->num)) return false;
  *out = (int64_t)v->num;
  return true;
}

// ------------------------------------------------------------------------------------------------ JSON (writer)
// shortest digits that round-trip, laid out like ryu's pretty printer (what serde_json emits): 0.5, 1.0, 1e-7, 1.5e300, -0.0
void put_f64(std::string* o, double x) {
  if (!std::isfinite(x)) { *o += "null"; return; }   // serde_json writes non-finite floats as null
  if (x == 0.0) { *o += std::signbit(x) ? "-0.0" : "0.0"; return; }
  char buf[40];
  int prec = 0;
  for (; prec < 17; ++prec) {
    snprintf(buf, sizeof(buf), "%.*e", prec, x);
    if (strtod(buf, nullptr) == x) break;
  }
  // buf = [-]d[.ddd]e[+-]XX
  std::string s(buf);
  const bool neg = s[0] == '-';
  if (neg) s.erase(0, 1);
  const size_t epos = s.find('e');
  std::string digits = s.substr(0, epos);
  const int exp10 = atoi(s.c_str() + epos + 1);
  digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
  const int n = (int)digits.size();
  const int kk = exp10 + 1;   // position of the decimal point relative to the first digit
  std::string out;
  if (neg) out += '-';
  if (n <= kk && kk <= 16) {               // 1234.0
    out += digits; out.append((size_t)(kk - n), '0'); out += ".0";
  } else if (0 < kk && kk <= 16) {         // 12.34
    out += digits.substr(0, (size_t)kk); out += '.'; out += digits.substr((size_t)kk);
  } else if (-5 < kk && kk <= 0) {         // 0.001234
    out += "0."; out.append((size_t)(-kk), '0'); out += digits;
  } else {                                 // 1.234e-7 / 1e21
    out += digits[0];
    if (n > 1) { out += '.'; out += digits.substr(1); }
    out += 'e'; out += std::to_string(kk - 1);
  }
  *o += out;
}
}  // namespace

struct porrt_graph {
  std::vector<double> xy;
  std::vector<int32_t> node_vid;
  std::vector<int64_t> row_ptr, p_row_ptr;
  std::vector<int32_t> col, edge_vid, p_col, p_edge_vid;
  std::vector<uint8_t> validities;   // [n_validities][n_worlds]
  int32_t n_validities = 0, n_worlds = 0;
};

// Map::open_image's decode step.  out == NULL or cap < h * w: PORRT_ERR_CAPACITY with *out_h / *out_w set (size query).
PORRT_API int32_t porrt_pgm_read(porrt_ctx* ctx, const char* path, uint8_t* out, int64_t cap, int32_t* out_h, int32_t* out_w) {
  if (!path || !out_h || !out_w) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "pgm_read: bad arguments");
  std::string d;
  if (!read_file(path, &d)) return fmt_fail(ctx, PORRT_ERR_PANIC, std::string("Impossible to open image: ") + path);   // map_io.rs:99
  size_t pos = 0;
  std::string magic, tw, th, tm;
  long w = 0, h = 0, maxval = 0;
  if (!pnm_token(d, &pos, &magic) || !pnm_token(d, &pos, &tw) || !pnm_token(d, &pos, &th) || !pnm_token(d, &pos, &tm) || !to_int(tw, &w) ||
      !to_int(th, &h) || !to_int(tm, &maxval) || w <= 0 || h <= 0 || w > 1000000 || h > 1000000)
    return fmt_fail(ctx, PORRT_ERR_PANIC, std::string("Impossible to open image: ") + path + " (malformed PNM header)");
  if ((magic != "P2" && magic != "P5") || maxval <= 0 || maxval >= 256)
    return fmt_fail(ctx, PORRT_ERR_PANIC, "Wrong image format! (only 8-bit P2 / P5 gray maps decode to ImageLuma8, map_io.rs:101-104)");
  *out_h = (int32_t)h; *out_w = (int32_t)w;
  const int64_t n = (int64_t)w * h;
  if (!out || cap < n) return fmt_fail(ctx, PORRT_ERR_CAPACITY, "pgm_read: output buffer too small");
  if (magic == "P5") {
    ++pos;   // exactly one white-space byte after maxval
    if (pos + (size_t)n > d.size()) return fmt_fail(ctx, PORRT_ERR_PANIC, std::string("Impossible to open image: ") + path + " (truncated)");
    memcpy(out, d.data() + pos, (size_t)n);
  } else {
    for (int64_t k = 0; k < n; ++k) {
      std::string t;
      long v = 0;
      if (!pnm_token(d, &pos, &t) || !to_int(t, &v) || v < 0 || v > maxval)
        return fmt_fail(ctx, PORRT_ERR_PANIC, std::string("Impossible to open image: ") + path + " (bad sample)");
      out[k] = (uint8_t)v;
    }
  }
  return PORRT_OK;
}

PORRT_API int32_t porrt_pgm_write(porrt_ctx* ctx, const char* path, const uint8_t* img, int32_t h, int32_t w, int32_t binary) {
  if (!path || !img || h <= 0 || w <= 0) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "pgm_write: bad arguments");
  std::ofstream f(path, std::ios::binary);
  if (!f) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, std::string("pgm_write: cannot create ") + path);
  f << (binary ? "P5\n" : "P2\n") << w << " " << h << "\n255\n";
  if (binary) f.write((const char*)img, (std::streamsize)h * w);
  else
    for (int i = 0; i < h; ++i) {
      for (int j = 0; j < w; ++j) f << (j ? " " : "") << (int)img[(size_t)i * w + j];
      f << "\n";
    }
  return f.good() ? PORRT_OK : fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "pgm_write: write failed");
}

// pto_graph::load (pto_graph.rs:110-118).  The handle owns the arrays; read them with porrt_graph_info / porrt_graph_arrays.
PORRT_API int32_t porrt_graph_load_json(porrt_ctx* ctx, const char* path, porrt_graph** out_graph) {
  if (!path || !out_graph) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "graph_load_json: bad arguments");
  *out_graph = nullptr;
  std::string d;
  if (!read_file(path, &d)) return fmt_fail(ctx, PORRT_ERR_PANIC, "impossible to open file (pto_graph.rs:111)");
  JParser jp(d);
  JVal root;
  if (!jp.value(&root)) return fmt_fail(ctx, PORRT_ERR_PANIC, "graph_load_json: " + jp.err + " (serde_json::from_reader(..).unwrap(), pto_graph.rs:112)");
  const JVal* nodes = root.get("nodes");
  const JVal* vals = root.get("validities");
  if (root.kind != JVal::OBJ || !nodes || nodes->kind != JVal::ARR || !vals || vals->kind != JVal::ARR)
    return fmt_fail(ctx, PORRT_ERR_PANIC, "graph_load_json: missing field `nodes` / `validities`");
  std::unique_ptr<porrt_graph> g(new porrt_graph());
  const int64_t V = (int64_t)nodes->arr.size();
  g->row_ptr.push_back(0); g->p_row_ptr.push_back(0);
  for (const JVal& n : nodes->arr) {
    const JVal* st = n.get("state");
    int64_t vid = 0;
    if (n.kind != JVal::OBJ || !st || st->kind != JVal::ARR || !as_index(n.get("validity_id"), &vid))
      return fmt_fail(ctx, PORRT_ERR_PANIC, "graph_load_json: a node lacks `state` / `validity_id`");
    if (st->arr.size() != 2 || st->arr[0].kind != JVal::NUM || st->arr[1].kind != JVal::NUM)
      return fmt_fail(ctx, PORRT_ERR_PANIC, "graph_load_json: state is not [f64; 2] (try_into().unwrap(), pto_graph.rs:66)");
    g->xy.push_back(st->arr[0].num); g->xy.push_back(st->arr[1].num);
    g->node_vid.push_back((int32_t)vid);
    for (int side = 0; side < 2; ++side) {
      const JVal* list = n.get(side == 0 ? "children" : "parents");
      if (!list || list->kind != JVal::ARR) return fmt_fail(ctx, PORRT_ERR_PANIC, "graph_load_json: a node lacks `children` / `parents`");
      std::vector<int32_t>& col = side == 0 ? g->col : g->p_col;
      std::vector<int32_t>& ev = side == 0 ? g->edge_vid : g->p_edge_vid;
      for (const JVal& e : list->arr) {
        int64_t id = 0, evid = 0;
        if (e.kind != JVal::OBJ || !as_index(e.get("id"), &id) || !as_index(e.get("validity_id"), &evid))
          return fmt_fail(ctx, PORRT_ERR_PANIC, "graph_load_json: an edge lacks `id` / `validity_id`");
        if (id >= V) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "graph_load_json: edge id out of range");
        col.push_back((int32_t)id); ev.push_back((int32_t)evid);
      }
      (side == 0 ? g->row_ptr : g->p_row_ptr).push_back((int64_t)col.size());
    }
  }
  g->n_validities = (int32_t)vals->arr.size();
  g->n_worlds = g->n_validities ? (int32_t)vals->arr[0].arr.size() : 0;
  for (const JVal& v : vals->arr) {
    if (v.kind != JVal::ARR || (int32_t)v.arr.size() != g->n_worlds) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "graph_load_json: validities of different lengths");
    for (const JVal& b : v.arr) {
      if (b.kind != JVal::BOOL) return fmt_fail(ctx, PORRT_ERR_PANIC, "graph_load_json: validities must be booleans");
      g->validities.push_back(b.b ? 1 : 0);
    }
  }
  *out_graph = g.release();
  return PORRT_OK;
}

PORRT_API int32_t porrt_graph_info(const porrt_graph* g, int64_t* out_n_nodes, int64_t* out_n_children, int64_t* out_n_parents,
                                   int32_t* out_n_validities, int32_t* out_n_worlds) {
  if (!g) return PORRT_ERR_INVALID_ARG;
  if (out_n_nodes) *out_n_nodes = (int64_t)g->node_vid.size();
  if (out_n_children) *out_n_children = (int64_t)g->col.size();
  if (out_n_parents) *out_n_parents = (int64_t)g->p_col.size();
  if (out_n_validities) *out_n_validities = g->n_validities;
  if (out_n_worlds) *out_n_worlds = g->n_worlds;
  return PORRT_OK;
}

// copies out whatever is asked for (every pointer nullable); sizes as reported by porrt_graph_info
PORRT_API int32_t porrt_graph_arrays(const porrt_graph* g, double* out_xy, int32_t* out_node_vid, int64_t* out_row_ptr, int32_t* out_col,
                                     int32_t* out_edge_vid, int64_t* out_p_row_ptr, int32_t* out_p_col, int32_t* out_p_edge_vid,
                                     uint8_t* out_validities) {
  if (!g) return PORRT_ERR_INVALID_ARG;
  auto put = [](void* dst, const void* src, size_t bytes) { if (dst && bytes) memcpy(dst, src, bytes); };
  put(out_xy, g->xy.data(), g->xy.size() * 8);
  put(out_node_vid, g->node_vid.data(), g->node_vid.size() * 4);
  put(out_row_ptr, g->row_ptr.data(), g->row_ptr.size() * 8);
  put(out_col, g->col.data(), g->col.size() * 4);
  put(out_edge_vid, g->edge_vid.data(), g->edge_vid.size() * 4);
  put(out_p_row_ptr, g->p_row_ptr.data(), g->p_row_ptr.size() * 8);
  put(out_p_col, g->p_col.data(), g->p_col.size() * 4);
  put(out_p_edge_vid, g->p_edge_vid.data(), g->p_edge_vid.size() * 4);
  put(out_validities, g->validities.data(), g->validities.size());
  return PORRT_OK;
}

PORRT_API int32_t porrt_graph_destroy(porrt_graph* g) {
  delete g;
  return PORRT_OK;
}

// pto_graph::save (pto_graph.rs:105-108): serde_json::to_writer_pretty layout (2-space indent, keys in struct order: state,
// validity_id, parents, children; nodes, validities).  p_* (parents) may be NULL: they are then derived from the children --
// parents(v) lists the sources of the edges u -> v in the order those edges appear when the rows are walked in node order, which is
// add_edge order for every graph whose edges were added row by row (a PRM, where parents(k) == children(k), prm.rs:99-106).
PORRT_API int32_t porrt_graph_save_json(porrt_ctx* ctx, const char* path, int64_t V, const double* xy, const int32_t* node_vid,
                                        const int64_t* row_ptr, const int32_t* col, const int32_t* edge_vid, const int64_t* p_row_ptr,
                                        const int32_t* p_col, const int32_t* p_edge_vid, const uint8_t* validities, int32_t n_validities,
                                        int32_t n_worlds) {
  if (!path || V < 0 || (V > 0 && (!xy || !node_vid || !row_ptr)) || n_validities < 0 || n_worlds < 0 || (n_validities * n_worlds > 0 && !validities))
    return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "graph_save_json: bad arguments");
  const int64_t E = V > 0 ? row_ptr[V] : 0;
  if (E > 0 && (!col || !edge_vid)) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "graph_save_json: null col / edge_vid");
  std::vector<int64_t> t_row;
  std::vector<int32_t> t_col, t_ev;
  if (!p_row_ptr) {
    t_row.assign((size_t)V + 1, 0);
    for (int64_t e = 0; e < E; ++e) {
      if (col[e] < 0 || col[e] >= V) return fmt_fail(ctx, PORRT_ERR_INVALID_ARG, "graph_save_json: edge id out of range");
      ++t_row[(size_t)col[e] + 1];
    }
    for (int64_t v = 0; v < V; ++v) t_row[(size_t)v + 1] += t_row[(size_t)v];
    t_col.resize((size_t)E); t_ev.resize((size_t)E);
    std::vector<int64_t> cur(t_row.begin(), t_row.end() - 1);
    for (int64_t u = 0; u < V; ++u)
      for (int64_t e = row_ptr[u]; e < row_ptr[u + 1]; ++e) { const int64_t at = cur[(size_t)col[e]]++; t_col[(size_t)at] = (int32_t)u; t_ev[(size_t)at] = edge_vid[e]; }
    p_row_ptr = t_row.data(); p_col = t_col.data(); p_edge_vid = t_ev.data();
  }
  std::string o;
  o.reserve((size_t)(V * 120 + E * 110 + 256));
  auto edges = [&](const char* key, const int64_t* rp, const int32_t* c, const int32_t* ev, int64_t k, bool last) {
    o += "      \""; o += key; o += "\": [";
    for (int64_t e = rp[k]; e < rp[k + 1]; ++e) {
      o += e == rp[k] ? "\n" : ",\n";
      o += "        {\n          \"id\": " + std::to_string(c[e]) + ",\n          \"validity_id\": " + std::to_string(ev[e]) + "\n        }";
    }
    o += rp[k + 1] > rp[k] ? "\n      ]" : "]";
    o += last ? "\n" : ",\n";
  };
  o += "{\n  \"nodes\": [";
  for (int64_t k = 0; k < V; ++k) {
    o += k ? ",\n" : "\n";
    o += "    {\n      \"state\": [\n        ";
    put_f64(&o, xy[2 * k]);
    o += ",\n        ";
    put_f64(&o, xy[2 * k + 1]);
    o += "\n      ],\n      \"validity_id\": " + std::to_string(node_vid[k]) + ",\n";
    edges("parents", p_row_ptr, p_col, p_edge_vid, k, false);
    edges("children", row_ptr, col, edge_vid, k, true);
    o += "    }";
  }
  o += V ? "\n  ],\n" : "],\n";
  o += "  \"validities\": [";
  for (int v = 0; v < n_validities; ++v) {
    o += v ? ",\n" : "\n";
    o += "    [";
    for (int w = 0; w < n_worlds; ++w) {
      o += w ? ",\n" : "\n";
      o += validities[(size_t)v * n_worlds + w] ? "      true" : "      false";
    }
    o += n_worlds ? "\n    ]" : "]";
  }
  o += n_validities ? "\n  ]\n}" : "]\n}";
  std::ofstream f(path, std::ios::binary);
  if (!f) return fmt_fail(ctx, PORRT_ERR_PANIC, "can't create file (pto_graph.rs:88)");
  f.write(o.data(), (std::streamsize)o.size());
  return f.good() ? PORRT_OK : fmt_fail(ctx, PORRT_ERR_PANIC, "error happened while dumping pto graph to file (pto_graph.rs:89)");
}

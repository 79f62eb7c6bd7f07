// mini_json.cpp -- see mini_json.h.  Independent implementation; the behaviours that matter are
// pinned against the reference's vendored jsoncpp by tests/golden/geometry_*.json.
#include "mini_json.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace mjson {

static const Value kNull;

void Value::reset() {
  type_ = Null; b_ = false; i_ = 0; u_ = 0; d_ = 0; s_.clear(); arr_.clear(); keys_.clear();
}

size_t Value::size() const { return (type_ == Array || type_ == Object) ? arr_.size() : 0; }

int Value::asInt() const {
  const double lo = std::numeric_limits<int>::min(), hi = std::numeric_limits<int>::max();
  switch (type_) {
    case Int:
      if (i_ < (long long)lo || i_ > (long long)hi) throw std::runtime_error("LargestInt out of Int range");
      return (int)i_;
    case UInt:
      if (u_ > (unsigned long long)hi) throw std::runtime_error("LargestUInt out of Int range");
      return (int)u_;
    case Real:
      if (!(d_ >= lo && d_ <= hi)) throw std::runtime_error("double out of Int range");
      return (int)d_;                       // truncation toward zero
    case Null: return 0;
    case Bool: return b_ ? 1 : 0;
    default: break;
  }
  throw std::runtime_error("Value is not convertible to Int.");
}

double Value::asDouble() const {
  switch (type_) {
    case Int: return (double)i_;
    case UInt: return (double)u_;
    case Real: return d_;
    case Null: return 0.0;
    case Bool: return b_ ? 1.0 : 0.0;
    default: break;
  }
  throw std::runtime_error("Value is not convertible to double.");
}

float Value::asFloat() const {
  switch (type_) {
    case Int: return (float)i_;
    case UInt: return (float)u_;
    case Real: return (float)d_;
    case Null: return 0.0f;
    case Bool: return b_ ? 1.0f : 0.0f;
    default: break;
  }
  throw std::runtime_error("Value is not convertible to float.");
}

bool Value::asBool() const {
  switch (type_) {
    case Bool: return b_;
    case Null: return false;
    case Int: return i_ != 0;
    case UInt: return u_ != 0;
    case Real: return d_ != 0.0;
    default: break;
  }
  throw std::runtime_error("Value is not convertible to bool.");
}

std::string Value::asString() const {
  char buf[64];
  switch (type_) {
    case Null: return "";
    case String: return s_;
    case Bool: return b_ ? "true" : "false";
    case Int: snprintf(buf, sizeof buf, "%lld", i_); return buf;
    case UInt: snprintf(buf, sizeof buf, "%llu", u_); return buf;
    case Real: snprintf(buf, sizeof buf, "%.17g", d_); return buf;
    default: break;
  }
  throw std::runtime_error("Type is not convertible to string");
}

const Value& Value::get(const std::string& key, const Value& dflt) const {
  if (type_ == Null) return dflt;
  if (type_ != Object) throw std::runtime_error("in Json::Value::find(key, end, found): requires objectValue or nullValue");
  for (size_t k = 0; k < keys_.size(); ++k)
    if (keys_[k] == key) return arr_[k];
  return dflt;
}

bool Value::has(const std::string& key) const {
  if (type_ != Object) return false;
  for (const auto& k : keys_) if (k == key) return true;
  return false;
}

const Value& Value::at(long long index) const {
  if (type_ == Null) return kNull;
  if (type_ != Array) throw std::runtime_error("in Json::Value::operator[](ArrayIndex): requires arrayValue");
  if (index < 0 || (size_t)index >= arr_.size()) return kNull;
  return arr_[(size_t)index];
}

Value& Value::append() { arr_.emplace_back(); return arr_.back(); }

Value& Value::member(const std::string& key) {
  for (size_t k = 0; k < keys_.size(); ++k)
    if (keys_[k] == key) return arr_[k];
  keys_.push_back(key);
  arr_.emplace_back();
  return arr_.back();
}

// ---------------------------------------------------------------------------------------
class Parser {
 public:
  Parser(const std::string& t) : p_(t.data()), end_(t.data() + t.size()) {}
  bool run(Value& root, std::string* err) {
    bool ok = readValue(root);
    if (err) *err = err_;
    return ok;
  }

 private:
  enum Tok { EndOfStream, ObjBegin, ObjEnd, ArrBegin, ArrEnd, Str, Num, True, False, NullTok, Comma, Colon, Comment, Error };
  struct Token { Tok type; const char* a; const char* b; };
  const char *p_, *end_;
  std::string err_;

  bool fail(const char* m) { if (err_.empty()) err_ = m; return false; }
  int next() { return p_ == end_ ? 0 : (unsigned char)*p_++; }
  bool match(const char* s, int n) {
    if (end_ - p_ < n) return false;
    for (int k = 0; k < n; ++k) if (p_[k] != s[k]) return false;
    p_ += n;
    return true;
  }
  void readToken(Token& t) {
    while (p_ != end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\r' || *p_ == '\n')) ++p_;
    t.a = p_;
    int c = next();
    bool ok = true;
    switch (c) {
      case '{': t.type = ObjBegin; break;
      case '}': t.type = ObjEnd; break;
      case '[': t.type = ArrBegin; break;
      case ']': t.type = ArrEnd; break;
      case ',': t.type = Comma; break;
      case ':': t.type = Colon; break;
      case 0: t.type = EndOfStream; break;
      case '"': {
        t.type = Str;
        ok = false;
        while (p_ != end_) {
          int d = next();
          if (d == '\\') next();
          else if (d == '"') { ok = true; break; }
        }
        break;
      }
      case '/': {
        t.type = Comment;
        int d = next();
        if (d == '*') {
          ok = false;
          while (p_ != end_) { if (next() == '*' && p_ != end_ && *p_ == '/') { ++p_; ok = true; break; } }
        } else if (d == '/') {
          while (p_ != end_) { int e = next(); if (e == '\n') break; if (e == '\r') { if (p_ != end_ && *p_ == '\n') ++p_; break; } }
        } else ok = false;
        break;
      }
      case 't': t.type = True; ok = match("rue", 3); break;
      case 'f': t.type = False; ok = match("alse", 4); break;
      case 'n': t.type = NullTok; ok = match("ull", 3); break;
      default:
        if ((c >= '0' && c <= '9') || c == '-') {
          t.type = Num;
          while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_;
          if (p_ != end_ && *p_ == '.') { ++p_; while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_; }
          if (p_ != end_ && (*p_ == 'e' || *p_ == 'E')) {
            ++p_;
            if (p_ != end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_;
          }
        } else ok = false;
    }
    if (!ok) t.type = Error;
    t.b = p_;
  }
  void readTokenSkippingComments(Token& t) { do readToken(t); while (t.type == Comment); }
  // Reader::recoverFromError: token-level skip, not nesting aware
  bool recover(Tok until) {
    Token t;
    for (;;) { readToken(t); if (t.type == until || t.type == EndOfStream) break; }
    return false;
  }
  static bool decodeString(const Token& t, std::string& out) {
    out.clear();
    const char* c = t.a + 1; const char* e = t.b - 1;
    while (c < e) {
      char ch = *c++;
      if (ch == '"') break;
      if (ch != '\\') { out += ch; continue; }
      if (c == e) return false;
      char esc = *c++;
      switch (esc) {
        case '"': out += '"'; break; case '/': out += '/'; break; case '\\': out += '\\'; break;
        case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
        case 'r': out += '\r'; break; case 't': out += '\t'; break;
        case 'u': {
          if (e - c < 4) return false;
          unsigned cp = 0;
          for (int k = 0; k < 4; ++k) {
            char h = *c++; cp <<= 4;
            if (h >= '0' && h <= '9') cp += h - '0'; else if (h >= 'a' && h <= 'f') cp += h - 'a' + 10;
            else if (h >= 'A' && h <= 'F') cp += h - 'A' + 10; else return false;
          }
          if (cp < 0x80) out += (char)cp;
          else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
          else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
          break;
        }
        default: return false;
      }
    }
    return true;
  }
  bool decodeNumber(const Token& t, Value& v) {
    const char* c = t.a;
    bool neg = *c == '-';
    if (neg) ++c;
    const unsigned long long maxv = neg ? (unsigned long long)std::numeric_limits<long long>::max() + 1ull
                                        : std::numeric_limits<unsigned long long>::max();
    const unsigned long long thr = maxv / 10;
    unsigned long long val = 0;
    bool as_double = false;
    while (c < t.b) {
      char ch = *c++;
      if (ch < '0' || ch > '9') { as_double = true; break; }
      unsigned d = ch - '0';
      if (val >= thr && (val > thr || c != t.b || d > maxv % 10)) { as_double = true; break; }
      val = val * 10 + d;
    }
    if (as_double) {
      std::string buf(t.a, t.b);
      std::istringstream is(buf);
      double d = 0;
      if (!(is >> d)) return fail("not a number");
      v = Value(d);
      return true;
    }
    if (neg && val == maxv) v = Value(std::numeric_limits<long long>::min());
    else if (neg) v = Value(-(long long)val);
    else if (val <= (unsigned long long)std::numeric_limits<int>::max()) v = Value((long long)val);
    else v = Value(val);
    return true;
  }
  bool readValue(Value& v) {
    Token t;
    readTokenSkippingComments(t);
    switch (t.type) {
      case ObjBegin: return readObject(v);
      case ArrBegin: return readArray(v);
      case Num: return decodeNumber(t, v);
      case Str: { std::string s; if (!decodeString(t, s)) return fail("bad string"); v = Value(s); return true; }
      case True: v = Value(true); return true;
      case False: v = Value(false); return true;
      case NullTok: v = Value(); return true;
      default: return fail("Syntax error: value, object or array expected.");
    }
  }
  bool readObject(Value& v) {
    v.makeObject();
    Token name;
    std::string key;
    for (;;) {
      readTokenSkippingComments(name);
      if (name.type == ObjEnd && key.empty()) return true;   // empty object
      key.clear();
      if (name.type != Str) break;
      if (!decodeString(name, key)) return recover(ObjEnd);
      Token colon;
      readToken(colon);
      if (colon.type != Colon) { fail("Missing ':' after object member name"); return recover(ObjEnd); }
      if (!readValue(v.member(key))) return recover(ObjEnd);
      Token comma;
      readTokenSkippingComments(comma);
      if (comma.type == ObjEnd) return true;
      if (comma.type != Comma) { fail("Missing ',' or '}' in object declaration"); return recover(ObjEnd); }
    }
    fail("Missing '}' or object member name");
    return recover(ObjEnd);
  }
  bool readArray(Value& v) {
    v.makeArray();
    while (p_ != end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\r' || *p_ == '\n')) ++p_;
    if (p_ != end_ && *p_ == ']') { ++p_; return true; }
    for (;;) {
      if (!readValue(v.append())) return recover(ArrEnd);
      Token t;
      readTokenSkippingComments(t);
      if (t.type == ArrEnd) return true;
      if (t.type != Comma) { fail("Missing ',' or ']' in array declaration"); return recover(ArrEnd); }
    }
  }
};

bool parse(const std::string& text, Value& root, std::string* error) {
  root.reset();
  Parser ps(text);
  return ps.run(root, error);
}

bool parseFile(const std::string& path, Value& root, std::string* error) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { root.reset(); if (error) *error = "cannot open " + path; return false; }
  std::stringstream ss;
  ss << f.rdbuf();
  return parse(ss.str(), root, error);
}

}  // namespace mjson

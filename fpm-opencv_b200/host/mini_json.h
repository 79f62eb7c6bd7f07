// mini_json.h -- small JSON DOM for the dataset*.json / ledArrayMaps formats.
//
// The reference reads its configuration with jsoncpp 1.6.5 (`Json::Reader`, `Json::Value`,
// fpmMain.cpp:512-584).  This is an independent reader that keeps the behaviours the shipped
// files depend on:
//   * `get(key, default)` with typed defaults; `asInt()` truncates reals; `asFloat()` casts;
//     accessors on a wrong type throw std::runtime_error (jsoncpp: Json::LogicError);
//   * numbers without '.', 'e', 'E' are integers, everything else goes through the C locale
//     string->double conversion;
//   * `//` and `/* */` comments are skipped;
//   * a trailing comma before ']' leaves one extra null element, reports failure and discards
//     the rest of the enclosing scopes token-by-token exactly like `Reader::recoverFromError`
//     (include/jsoncpp.cpp:941-953) -- dataset_cellScope.json:538-539 and
//     dataset_dogStomach.json:320-321 end that way and the reference ignores the failure
//     (fpmMain.cpp:515);
//   * indexing a null / short array yields null (the reference indexes through the non-const
//     `operator[]`, fpmMain.cpp:77-79, which auto-creates nulls).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mjson {

class Value {
 public:
  enum Type { Null, Int, UInt, Real, String, Bool, Array, Object };
  Value() = default;
  explicit Value(bool b) : type_(Bool), b_(b) {}
  explicit Value(int i) : type_(Int), i_(i) {}
  explicit Value(long long i) : type_(Int), i_(i) {}
  explicit Value(unsigned long long u) : type_(UInt), u_(u) {}
  explicit Value(double d) : type_(Real), d_(d) {}
  explicit Value(const char* s) : type_(String), s_(s) {}
  explicit Value(const std::string& s) : type_(String), s_(s) {}

  Type type() const { return type_; }
  bool isNull() const { return type_ == Null; }
  bool isArray() const { return type_ == Array; }
  bool isObject() const { return type_ == Object; }
  bool isNumeric() const { return type_ == Int || type_ == UInt || type_ == Real; }
  size_t size() const;

  int asInt() const;            // Value::asInt, include/jsoncpp.cpp:3048-3069
  double asDouble() const;      // :3153-3173
  float asFloat() const;        // :3175-3195
  bool asBool() const;          // :3197-3215
  std::string asString() const; // :3013-3036

  // object member or `dflt` (Value::get).  Throws on non-object, non-null values.
  const Value& get(const std::string& key, const Value& dflt) const;
  bool has(const std::string& key) const;
  // array element or null (mirrors auto-creating operator[] on a copy); throws on scalars
  const Value& at(long long index) const;

  // builders used by the parser
  void makeArray() { reset(); type_ = Array; }
  void makeObject() { reset(); type_ = Object; }
  Value& append();                        // new null element
  Value& member(const std::string& key);  // create / fetch member
  void reset();

 private:
  Type type_ = Null;
  bool b_ = false;
  long long i_ = 0;
  unsigned long long u_ = 0;
  double d_ = 0;
  std::string s_;
  std::vector<Value> arr_;              // array elements, or object member values
  std::vector<std::string> keys_;       // object member names (parallel to arr_)
  friend class Parser;
};

// Parses `text` into `root`.  Returns what Json::Reader::parse would return; like the reference,
// callers may ignore it -- `root` keeps everything that was read before the first error.
bool parse(const std::string& text, Value& root, std::string* error = nullptr);
bool parseFile(const std::string& path, Value& root, std::string* error = nullptr);

}  // namespace mjson
